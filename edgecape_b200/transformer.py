"""@TRANSFORMER TwoStageSupportRefineTransformer on the edgecape_b200 kernels.

Same constructor kwargs, `forward` signature and state-dict keys as
/root/reference/EdgeCape/models/keypoint_heads/encoder_decoder.py:115-260, with the arithmetic
of TransformerEncoder(:268-310,434-483), ProposalGenerator (:37-112), TransformerDecoder
(:313-431), TransformerDecoderLayer (:527-651), GCNLayer (:486-524) and BiasedMultiheadAttention
(utils/bias_attn.py:106-231) executed by the C-ABI kernels.  Activations are batch-first
token-major [B, tokens, C]; the reference's [L, B, C] permutes are pure data movement and never
happen on the product path (`forward_tokens`).  Inference only: dropout is inert under eval()
and the masked-supervision branch (:212-237) is training-only.
"""
import torch
import torch.nn as nn

from . import ops
from .config import decoder_layer_shapes, _lin_keys, _ln_keys, _mha_keys
from .params import PackedMixin, ParamTree, xavier_uniform_all_
from .registry import TRANSFORMER


def pack_decoder_layer(node, biased, bias_mlp, two_way):
    """Kernel-native copies of one TransformerDecoderLayer's weights."""
    pk = {}
    sa = node.self_attn
    if biased:
        # q/k/v_proj fused into one [3d, d] GEMM (utils/bias_attn.py:149-155)
        pk["qkv_w"] = torch.cat((sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight)).contiguous()
        pk["qkv_b"] = torch.cat((sa.q_proj.bias, sa.k_proj.bias, sa.v_proj.bias)).contiguous()
        if bias_mlp:
            m = sa.markov_structural_mlp
            pk["hop"] = (getattr(m, "0").weight, getattr(m, "0").bias, getattr(m, "3").weight, getattr(m, "3").bias)
    else:
        pk["qkv_w"], pk["qkv_b"] = sa.in_proj_weight, sa.in_proj_bias
    pk["gcn_w"] = ops.gcn_pack_weights(node.ffn1.conv.weight, node.ffn1.conv.bias)
    for name in (["multihead_attn"] + (["cross_attn_image_to_token"] if two_way else [])):
        a = getattr(node, name)
        E = a.q_proj_weight.shape[0]
        b = a.in_proj_bias
        pk[name] = dict(qb=b[:E].contiguous(), kb=b[E:2 * E].contiguous(), vb=b[2 * E:].contiguous())
        # K and V projections as ONE GEMM over the [tokens | pos] operand: rows [0, E) = W_k, rows [E, 2E) =
        # [W_v | 0] (the value projection reads only the token half, encoder_decoder.py:622-628, 640-646)
        wv = torch.zeros_like(a.k_proj_weight)
        wv[:, :a.v_proj_weight.shape[1]] = a.v_proj_weight
        pk[name]["kv_w"] = torch.cat((a.k_proj_weight, wv)).contiguous()
        pk[name]["kv_b"] = b[E:].contiguous()
        # out_proj followed by the choker is one linear map (no non-linearity in between, :629-631 / :647-649):
        # W = W_choker W_out, b = W_choker b_out + b_choker, formed once with the fp32 GEMM kernel
        ch = node.choker if name == "multihead_attn" else node.cross_attn_image_to_token_choker
        pk[name]["oc_w"] = ops.gemm(ch.weight, a.out_proj.weight, b_kmajor=False)
        pk[name]["oc_b"] = ops.gemm(a.out_proj.bias.view(1, -1), ch.weight, b_kmajor=True, bias=ch.bias).view(-1)
    return pk


def _split_attention_ok(x2d, w, D, Lk, masked=False):
    """Tensor-core mode and shapes the split-fp16 path takes: the projections run on the tcgen05 GEMM with a
    split-fp16 epilogue and ec_attention_tc_split consumes them (no fp32 Q/K/V round trip)."""
    return ops.attention_split_ok(D, Lk, masked) and ops.tc_linear_ok(x2d, w)


def decoder_layer_forward(node, pk, nhead, kp, img_cat, kp_cat, key_mask_fixed, adj, attn_adj=None, two_way=False,
                          img_split=None, kv_split=None, kv_col=0, kp_split=None):
    """One TransformerDecoderLayer (encoder_decoder.py:584-651), batch-first.

    kp       [B,K,d]   keypoint tokens (contiguous)
    img_cat  [B,S,2d]  [:, :, :d] image tokens, [:, :, d:] their positional encoding (cat of :621)
    kp_cat   [B,K,2d]  scratch; [:, :, d:] already holds the keypoint positional embedding (:620)
    img_split          optional split-fp16 copy of img_cat (the caller keeps it across layers when the
                       image tokens do not change, i.e. two_way=False)
    kv_split, kv_col   optional precomputed [k | v] projection of the image tokens (split-fp16, this layer's
                       columns start at kv_col): the main decoder projects all its layers in one GEMM
    kp_split           optional split operand of kp (the previous layer's norm3 writes it next to the fp32 tokens)
    Returns (new keypoint tokens [B,K,d], their split operand or None); with two_way the image half of img_cat is
    updated in place (norm4 output feeds the next layer, :638-649)."""
    B, K, d = kp.shape
    S = img_cat.shape[1]
    dev = kp.device
    kp2d = kp.view(B * K, d)
    # (i) self-attention over keypoints (+ structural bias), residual, norm1
    bias = None
    use_hop = attn_adj is not None and "hop" in pk
    if _split_attention_ok(kp2d, pk["qkv_w"], d // nhead, K, masked=True):
        qkv2 = ops.linear_split(kp_split if kp_split is not None else kp2d, pk["qkv_w"], pk["qkv_b"])
        if use_hop and ops.HOP_FUSED and ops.hop_fused_ok(attn_adj.shape[0], pk["hop"][0].shape[0]):
            # the structural bias (utils/bias_attn.py:188-191) is evaluated inside the attention kernel, per logit
            a = ops.attention_packed_split(qkv2, B, K, nhead, key_mask=key_mask_fixed, hop=(attn_adj,) + tuple(pk["hop"]))
        else:
            if use_hop:
                bias = ops.hop_bias(attn_adj, *pk["hop"])
            a = ops.attention_packed_split(qkv2, B, K, nhead, key_mask=key_mask_fixed, bias=bias)
    else:
        if use_hop:
            bias = ops.hop_bias(attn_adj, *pk["hop"])
        qkv = ops.linear(kp2d, pk["qkv_w"], pk["qkv_b"]).view(B, K, 3 * d)
        a = ops.attention(qkv[:, :, 0:d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:], nhead, key_mask=key_mask_fixed,
                          bias=bias).view(B * K, d)
    t = ops.linear(a, node.self_attn.out_proj.weight, node.self_attn.out_proj.bias, residual=kp2d)
    kc2 = kp_cat.view(B * K, 2 * d)
    ops.layernorm(t, node.norm1.weight, node.norm1.bias, 1e-5, out=kc2[:, :d])
    kp1 = kc2[:, :d]
    # (ii) cross-attention: q = [kp | kp_pos], k = [img | pos], v = img  (8 heads x 64), choker
    ca, cp = node.multihead_attn, pk["multihead_attn"]
    ic2 = img_cat.view(B * S, 2 * d)
    Dc = 2 * d // nhead
    split_x = _split_attention_ok(kc2, ca.q_proj_weight, Dc, max(S, K)) and ops.tc_linear_ok(ic2, cp["kv_w"])
    if split_x:
        if img_split is None and (kv_split is None or two_way):
            img_split = ops.split_f16(ic2)
        q2 = ops.linear_split(kc2, ca.q_proj_weight, cp["qb"])
        if kv_split is None:
            kv_split, kv_col = ops.linear_split(img_split, cp["kv_w"], cp["kv_b"]), 0     # [B*S, 4d]: k | v
        a = ops.attention_split(q2, 0, K, kv_split, kv_col, kv_split, kv_col + 2 * d, S, B, nhead, K, S, Dc)
    else:
        q = ops.linear(kc2, ca.q_proj_weight, cp["qb"]).view(B, K, 2 * d)
        k = ops.linear(ic2, ca.k_proj_weight, cp["kb"]).view(B, S, 2 * d)
        v = ops.linear(ic2[:, :d], ca.v_proj_weight, cp["vb"]).view(B, S, 2 * d)
        a = ops.attention(q, k, v, nhead).view(B * K, 2 * d)
    if "oc_w" in cp:
        t = ops.linear(a, cp["oc_w"], cp["oc_b"], residual=kp1)
    else:
        a = ops.linear(a, ca.out_proj.weight, ca.out_proj.bias)
        t = ops.linear(a, node.choker.weight, node.choker.bias, residual=kp1)
    kp2 = ops.layernorm(t, node.norm2.weight, node.norm2.bias, 1e-5)
    # (iii) GCN feed-forward, ffn2, residual, norm3
    if ops.gcn_tc_ok(B, K):       # the GCN GEMM epilogue emits the split-fp16 operand of ffn2 directly
        g = ops.gcn(kp2.view(B, K, d), adj, pk["gcn_w"], split="only")
    else:
        g = ops.gcn(kp2.view(B, K, d), adj, pk["gcn_w"]).view(B * K, -1)
    t = ops.linear(g, node.ffn2.weight, node.ffn2.bias, residual=kp2)
    # norm3 also writes the split operand of its consumers (the next layer's q/k/v projection, the keypoint branch)
    want_split = ops.tc_linear_ok(kp2d, pk["qkv_w"])
    if not two_way:
        if want_split:
            y, ys = ops.layernorm(t, node.norm3.weight, node.norm3.bias, 1e-5, split="also")
            return y.view(B, K, d), ys
        return ops.layernorm(t, node.norm3.weight, node.norm3.bias, 1e-5).view(B, K, d), None
    # (iv) image tokens attend to the (un-masked!) keypoint tokens, choker, residual, norm4
    kp3_split = None
    if want_split:
        _, kp3_split = ops.layernorm(t, node.norm3.weight, node.norm3.bias, 1e-5, out=kc2[:, :d], split="also")
    else:
        ops.layernorm(t, node.norm3.weight, node.norm3.bias, 1e-5, out=kc2[:, :d])
    kp3 = kc2[:, :d]
    ia, ip = node.cross_attn_image_to_token, pk["cross_attn_image_to_token"]
    if split_x:
        q2 = ops.linear_split(img_split, ia.q_proj_weight, ip["qb"])                  # image tokens unchanged since (ii)
        kv2 = ops.linear_split(kc2, ip["kv_w"], ip["kv_b"])                           # [B*K, 4d]: k | v
        a = ops.attention_split(q2, 0, S, kv2, 0, kv2, 2 * d, K, B, nhead, S, K, Dc)
    else:
        q = ops.linear(ic2, ia.q_proj_weight, ip["qb"]).view(B, S, 2 * d)
        k = ops.linear(kc2, ia.k_proj_weight, ip["kb"]).view(B, K, 2 * d)
        v = ops.linear(kp3, ia.v_proj_weight, ip["vb"]).view(B, K, 2 * d)
        a = ops.attention(q, k, v, nhead).view(B * S, 2 * d)
    if "oc_w" in ip:
        t = ops.linear(a, ip["oc_w"], ip["oc_b"], residual=ic2[:, :d])
    else:
        a = ops.linear(a, ia.out_proj.weight, ia.out_proj.bias)
        t = ops.linear(a, node.cross_attn_image_to_token_choker.weight, node.cross_attn_image_to_token_choker.bias,
                       residual=ic2[:, :d])
    ops.layernorm(t, node.norm4.weight, node.norm4.bias, 1e-5, out=ic2[:, :d])
    out = ops.empty(B, K, d, device=dev)
    ops.copy_rows(kp3, out.view(B * K, d))
    return out, kp3_split


def _chain(x, lins, out=None):
    """Linear + GELU ... Linear.  On the tensor-core path the intermediate activations exist only as the split operand
    the next GEMM consumes (written by the producing GEMM's epilogue: no fp32 round trip, no conversion launch); `x`
    may itself be a SplitOperand."""
    n = len(lins)
    for i, lin in enumerate(lins):
        if i == n - 1:
            return ops.linear(x, lin.weight, lin.bias, out=out)
        if ops.tc_linear_ok(x, lin.weight):
            x = ops.linear(x, lin.weight, lin.bias, act=ops.ACT_GELU, split_out=True, fp32_out=False)[1]
        else:
            x = ops.linear(x, lin.weight, lin.bias, act=ops.ACT_GELU)
    return x


def mlp_gelu(x, node, n, out=None):
    """encoder_decoder.py:21-34 MLP (Linear+GELU ... Linear)."""
    return _chain(x, [getattr(node.layers, str(i)) for i in range(n)], out=out)


def token_decode_mlp(x, node):
    """head.py:34-58 TokenDecodeMLP: 3 x (Linear + GELU) + Linear(->2); x [M,d] (or its split operand) -> [M,2]."""
    return _chain(x, [getattr(node.mlp, k) for k in ("0", "2", "4", "6")])


@TRANSFORMER.register_module(force=True)
class TwoStageSupportRefineTransformer(PackedMixin, nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=3, num_decoder_layers=3, dim_feedforward=2048,
                 dropout=0.1, activation="relu", normalize_before=False, similarity_proj_dim=256,
                 dynamic_proj_dim=128, return_intermediate_dec=True, attn_bias=False, max_hops=5,
                 use_bias_attn_module=False, masked_supervision=False, recon_features=False):
        super().__init__()
        self._init_packed()
        if normalize_before:
            raise NotImplementedError("normalize_before=True is not used by any EdgeCape config")
        if activation != "relu":
            raise NotImplementedError("only activation='relu' is used by the EdgeCape configs")
        self.d_model, self.nhead = d_model, nhead
        self.num_encoder_layers, self.num_decoder_layers = num_encoder_layers, num_decoder_layers
        self.dim_feedforward = dim_feedforward
        self.attn_bias = attn_bias
        self.biased = attn_bias or use_bias_attn_module
        self.max_hops = max_hops
        self.masked_supervision = masked_supervision
        self.recon_features = recon_features
        self.return_intermediate_dec = return_intermediate_dec
        self.freeze = ""
        d, dff = d_model, dim_feedforward
        enc = {}
        for i in range(num_encoder_layers):
            q = f"layers.{i}"
            enc.update(_mha_keys(q + ".self_attn", d))
            enc.update(_lin_keys(q + ".linear1", dff, d))
            enc.update(_lin_keys(q + ".linear2", d, dff))
            enc.update(_ln_keys(q + ".norm1", d))
            enc.update(_ln_keys(q + ".norm2", d))
        self.encoder = ParamTree(enc) if num_encoder_layers > 0 else None
        dec = {}
        for i in range(num_decoder_layers):
            dec.update(decoder_layer_shapes(f"layers.{i}", d, nhead, dff, self.biased, max_hops, False, attn_bias))
        dec.update(_ln_keys("norm", d))
        dec.update(_lin_keys("ref_point_head.layers.0", d, d))
        dec.update(_lin_keys("ref_point_head.layers.1", d, d))
        self.decoder = ParamTree(dec)
        pg = {}
        pg.update(_lin_keys("support_proj", similarity_proj_dim, d))
        pg.update(_lin_keys("query_proj", similarity_proj_dim, d))
        pg.update(_lin_keys("dynamic_proj.0", dynamic_proj_dim, d))
        pg.update(_lin_keys("dynamic_proj.2", d, dynamic_proj_dim))
        self.proposal_generator = ParamTree(pg)
        for n, p in self.named_parameters():
            if "norm" in n and n.endswith("weight"):
                nn.init.ones_(p)

    def init_weights(self):
        xavier_uniform_all_(self)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        """Stage-2 checkpoints carry fused `in_proj_{weight,bias}` for the decoder self-attention;
        the biased module splits them into q/k/v_proj (utils/bias_attn.py:236-265)."""
        if self.biased:
            for i in range(self.num_decoder_layers):
                p = f"{prefix}decoder.layers.{i}.self_attn."
                for kind in ("weight", "bias"):
                    key = p + "in_proj_" + kind
                    if key in state_dict:
                        t = state_dict.pop(key)
                        dim = t.shape[0] // 3
                        state_dict[p + "q_proj." + kind] = t[:dim]
                        state_dict[p + "k_proj." + kind] = t[dim:2 * dim]
                        state_dict[p + "v_proj." + kind] = t[2 * dim:]
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _pack(self):
        dec = [pack_decoder_layer(getattr(self.decoder.layers, str(i)), self.biased, self.attn_bias, False)
               for i in range(self.num_decoder_layers)]
        # the image tokens are the same for every decoder layer: their K | V projections are one GEMM
        kv_w = torch.cat([pk["multihead_attn"]["kv_w"] for pk in dec]).contiguous()
        kv_b = torch.cat([pk["multihead_attn"]["kv_b"] for pk in dec]).contiguous()
        return {"dec": dec, "kv_w": kv_w, "kv_b": kv_b}

    # --------------------------------------------------------------------------- pieces
    def encode(self, x, grid_pos, S, key_mask):
        """3 x TransformerEncoderLayer (:461-483) over x = [B, S+K, d] (image tokens then keypoint
        tokens), in place.  grid_pos [S,d] is added to the residual stream at every layer."""
        B, T, d = x.shape
        x2 = x.view(B * T, d)
        for i in range(self.num_encoder_layers):
            L = getattr(self.encoder.layers, str(i))
            ops.add_rows_(x, grid_pos, S)
            if _split_attention_ok(x2, L.self_attn.in_proj_weight, d // self.nhead, T, masked=True):
                qkv2 = ops.linear_split(x2, L.self_attn.in_proj_weight, L.self_attn.in_proj_bias)
                a = ops.attention_packed_split(qkv2, B, T, self.nhead, key_mask=key_mask)
            else:
                qkv = ops.linear(x2, L.self_attn.in_proj_weight, L.self_attn.in_proj_bias).view(B, T, 3 * d)
                a = ops.attention(qkv[:, :, 0:d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:], self.nhead,
                                  key_mask=key_mask).view(B * T, d)
            t = ops.linear(a, L.self_attn.out_proj.weight, L.self_attn.out_proj.bias, residual=x2)
            if ops.tc_linear_ok(x2, L.linear1.weight):
                # norm1 also writes the split operand of linear1 (in the row format that GEMM takes), linear1 writes
                # nothing but the split operand of linear2: two conversion launches and an fp32 round trip less
                fmt = ops.F16F8 if ops.f8_linear_ok(B * T, L.linear1.weight) else ops.F16X2
                _, x2s = ops.layernorm(t, L.norm1.weight, L.norm1.bias, 1e-5, out=x2, split="also", split_fmt=fmt)
                fmt2 = ops.F16F8 if fmt == ops.F16F8 and ops.f8_linear_ok(B * T, L.linear2.weight) else ops.F16X2
                f = ops.linear(x2s, L.linear1.weight, L.linear1.bias, act=ops.ACT_RELU, split_out=True, fp32_out=False,
                               split_fmt=fmt2)[1]
            else:
                ops.layernorm(t, L.norm1.weight, L.norm1.bias, 1e-5, out=x2)
                f = ops.linear(x2, L.linear1.weight, L.linear1.bias, act=ops.ACT_RELU)
            t = ops.linear(f, L.linear2.weight, L.linear2.bias, residual=x2)
            ops.layernorm(t, L.norm2.weight, L.norm2.bias, 1e-5, out=x2)
        return x

    def propose(self, img, kp, h, w):
        """ProposalGenerator (:49-112): img [B,S,d] view, kp [B,K,d] view ->
        (proposal_for_loss [B,K,2], similarity [B,K,h,w], proposals [B,K,2], argmax [B,K])."""
        pg = self.proposal_generator
        B, S, _ = img.shape
        K = kp.shape[1]
        fs = ops.linear(kp, pg.support_proj.weight, pg.support_proj.bias)               # [B,K,p]
        fq = ops.linear(img, pg.query_proj.weight, pg.query_proj.bias)                  # [B,S,p]
        d0, d2 = getattr(pg.dynamic_proj, "0"), getattr(pg.dynamic_proj, "2")
        if ops.tc_linear_ok(fs, d0.weight):        # the hidden layer only exists as the split operand of the next GEMM
            hdn = ops.linear(fs, d0.weight, d0.bias, act=ops.ACT_RELU, split_out=True, fp32_out=False)[1]
        else:
            hdn = ops.linear(fs, d0.weight, d0.bias, act=ops.ACT_RELU)
        fsf = ops.linear(hdn, d2.weight, d2.bias, act=ops.ACT_TANH, residual=fs, res_mode=ops.RES_GATE)   # (tanh(.)+1) * fs
        fsf = fsf.view(B, K, -1)
        sim = ops.gemm(fsf, fq, b_kmajor=True)                                           # [B,K,S] exact fp32
        pl, pr, am = ops.proposal(sim, h, w)
        return pl, sim.view(B, K, h, w), pr, am

    # ----------------------------------------------------------------------------- main
    @torch.no_grad()
    def forward_tokens(self, x, S, hw, grid_pos, kp_mask, kp_mask_fixed, position_embedding, kpt_branch, adj,
                       attn_adj):
        """x [B, S+K, d]: image tokens (input_proj output) followed by support keypoint tokens.
        Returns dict(hs [L,B,K,d], out_points list of L+1 [B,K,2], proposal_for_loss, similarity_map,
        proposals, argmax)."""
        B, T, d = x.shape
        K = T - S
        h, w = hw
        dev = x.device
        enc_mask = ops.empty(B, T, dtype=torch.uint8, device=dev)
        enc_mask[:, :S].zero_()
        enc_mask[:, S:].copy_(kp_mask)
        if self.encoder is not None:
            self.encode(x, grid_pos, S, enc_mask)
        img, kp = x[:, :S, :], x[:, S:, :]
        pl, sim, pr, am = self.propose(img, kp, h, w)
        # decoder (:330-425).  `adj` may be a callable: the head runs the skeleton predictor on a side stream
        # concurrently with the encoder above and joins here, where its result is first needed
        if callable(adj):
            adj, attn_adj = adj()
        packed = self.packed()
        pk = packed["dec"]
        img_cat = ops.empty(B, S, 2 * d, device=dev)                 # [img | grid pos]  (torch.cat of :621)
        ops.copy_rows(img, img_cat[:, :, :d])
        ops.copy_rows(grid_pos, img_cat.view(B * S, 2 * d)[:, d:], bcast_rows=S)
        kp_cat = ops.empty(B, K, 2 * d, device=dev)
        cur = ops.copy_rows(kp, ops.empty(B, K, d, device=dev))
        bi = pr
        points = [pr]
        hs = ops.empty(self.num_decoder_layers, B, K, d, device=dev)
        use_bias = self.attn_bias and attn_adj is not None
        kv_all = None
        cur_split = None
        hs_split = []
        for i in range(self.num_decoder_layers):
            L = getattr(self.decoder.layers, str(i))
            pe = position_embedding.forward_coordinates(bi)                               # [B,K,d]
            mlp_gelu(pe.view(B * K, d), self.decoder.ref_point_head, 2, out=kp_cat.view(B * K, 2 * d)[:, d:])
            if i == 0 and ops.tc_linear_ok(img_cat.view(B * S, 2 * d), packed["kv_w"]) and \
                    ops.attention_split_ok(2 * d // self.nhead, max(S, K)):
                # image tokens are fixed over the layers: one split, one [k | v] projection GEMM for all of them
                kv_all = ops.linear_split(ops.split_f16(img_cat.view(B * S, 2 * d)), packed["kv_w"], packed["kv_b"])
            cur, cur_split = decoder_layer_forward(L, pk[i], self.nhead, cur, img_cat, kp_cat, kp_mask_fixed, adj,
                                                   attn_adj if use_bias else None, two_way=False, kv_split=kv_all,
                                                   kv_col=i * 4 * d, kp_split=cur_split)
            if cur_split is not None:      # the head's final decode runs the keypoint branch on hs[i]: split operand too
                hs_split.append(ops.layernorm(cur.view(B * K, d), self.decoder.norm.weight, self.decoder.norm.bias, 1e-5,
                                              out=hs[i].view(B * K, d), split="also")[1])
            else:
                hs_split.append(None)
                ops.layernorm(cur.view(B * K, d), self.decoder.norm.weight, self.decoder.norm.bias, 1e-5,
                              out=hs[i].view(B * K, d))
            delta = token_decode_mlp(cur_split if cur_split is not None else cur.view(B * K, d),
                                     getattr(kpt_branch, str(i)))
            bi = ops.point_update(bi, delta)
            points.append(bi)
        return dict(hs=hs, hs_split=hs_split, out_points=points, proposal_for_loss=pl, similarity_map=sim, proposals=pr, argmax=am,
                    encoder_image=img, encoder_kp=kp)

    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            "TwoStageSupportRefineTransformer is driven through TwoStageHead.forward / forward_tokens in this "
            "implementation (batch-first token-major tensors)")
