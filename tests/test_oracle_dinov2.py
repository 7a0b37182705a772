"""CPU: the DINOv2 restatement (oracle/dinov2_oracle.py) against an independent implementation of
the same architecture, `transformers.Dinov2Model`, with shared random weights.  DINOv2 is not
vendored in the reference (torch.hub, unpinned) and no upstream weights exist offline, so this is
the pin for that boundary."""
import pytest
import torch

from oracle.dinov2_oracle import vit_forward_tokens, vit_param_shapes

transformers = pytest.importorskip("transformers")


def _hf_to_upstream(hf_sd, depth):
    m = {"cls_token": "embeddings.cls_token", "mask_token": "embeddings.mask_token",
         "pos_embed": "embeddings.position_embeddings",
         "patch_embed.proj.weight": "embeddings.patch_embeddings.projection.weight",
         "patch_embed.proj.bias": "embeddings.patch_embeddings.projection.bias",
         "norm.weight": "layernorm.weight", "norm.bias": "layernorm.bias"}
    sd = {k: hf_sd[v] for k, v in m.items()}
    sd["mask_token"] = sd["mask_token"].reshape(1, -1)
    for i in range(depth):
        p, h = f"blocks.{i}.", f"encoder.layer.{i}."
        for n in ("norm1", "norm2"):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = hf_sd[h + n + ".weight"], hf_sd[h + n + ".bias"]
        a = h + "attention.attention."
        sd[p + "attn.qkv.weight"] = torch.cat([hf_sd[a + f"{n}.weight"] for n in ("query", "key", "value")])
        sd[p + "attn.qkv.bias"] = torch.cat([hf_sd[a + f"{n}.bias"] for n in ("query", "key", "value")])
        sd[p + "attn.proj.weight"] = hf_sd[h + "attention.output.dense.weight"]
        sd[p + "attn.proj.bias"] = hf_sd[h + "attention.output.dense.bias"]
        sd[p + "ls1.gamma"], sd[p + "ls2.gamma"] = hf_sd[h + "layer_scale1.lambda1"], hf_sd[h + "layer_scale2.lambda1"]
        for n in ("fc1", "fc2"):
            sd[p + f"mlp.{n}.weight"], sd[p + f"mlp.{n}.bias"] = hf_sd[h + f"mlp.{n}.weight"], hf_sd[h + f"mlp.{n}.bias"]
    return sd


@pytest.mark.parametrize("dim,depth,heads,grid", [(64, 2, 4, 4), (96, 3, 6, 5)])
def test_dinov2_restatement_matches_transformers(dim, depth, heads, grid):
    P = 14
    torch.manual_seed(0)
    hf_cfg = transformers.Dinov2Config(hidden_size=dim, num_hidden_layers=depth, num_attention_heads=heads,
                                       mlp_ratio=4, image_size=P * grid, patch_size=P, layerscale_value=1.0,
                                       hidden_act="gelu", layer_norm_eps=1e-6, qkv_bias=True)
    hf = transformers.Dinov2Model(hf_cfg).eval()
    with torch.no_grad():
        for p in hf.parameters():
            p.add_(0.05 * torch.randn_like(p))          # make every parameter (LayerScale, biases) non-trivial
    cfg = dict(embed_dim=dim, depth=depth, num_heads=heads, patch_size=P, img_size=P * grid, mlp_ratio=4,
               interpolate_offset=0.1)
    sd = _hf_to_upstream(hf.state_dict(), depth)
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v) for k, v in vit_param_shapes(cfg).items()}
    x = torch.randn(2, 3, P * grid, P * grid)
    with torch.no_grad():
        want = hf(pixel_values=x).last_hidden_state          # [B, 1+S, C], final LayerNorm applied
        got, (h0, w0) = vit_forward_tokens(sd, cfg, x)
    assert (h0, w0) == (grid, grid)
    err = (got - want[:, 1:]).abs().max().item() / want.abs().max().item()
    assert err < 2e-5, err
