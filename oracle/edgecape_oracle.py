"""TEST INFRASTRUCTURE ONLY -- CPU restatement (PyTorch, fp32 or fp64) of EdgeCape's
per-image inference hot path.  It is the *checker* for the CUDA path; only tests/,
bench.py's cpu_baseline / --impl reference leg and __graft_entry__.smoke() may import it.
The product (edgecape_b200/) never does, and fails loudly without its CUDA library.

PARITY PIN: every function below is checked against the reference's own modules
(imported unmodified from /root/reference through oracle/ref_shims.py) by
oracle/gen_golden.py, which also freezes the reference's outputs as tests/golden/*.npz;
tests/test_oracle_golden.py re-checks this restatement against those files wherever
the repo travels.  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so the goldens are outputs of the reference itself run here.

Each function cites the reference lines it restates (paths relative to
/root/reference/EdgeCape/models/).  Weights come in as a flat state dict with the
reference's key names.  Layout is batch-first token-major ([B, tokens, C]) throughout;
the reference's sequence-first permutes are pure data movement.
"""
import math

import torch
import torch.nn.functional as F

from .dinov2_oracle import vit_config, vit_forward_tokens


# ----------------------------------------------------------------------------- helpers
def inverse_sigmoid(x, eps=1e-3):
    """keypoint_heads/head.py:27-31 and encoder_decoder.py:14-18."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def sine_pe_grid(h, w, dtype, num_feats=128, temperature=10000, scale=2 * math.pi, eps=1e-6):
    """utils/positional_encoding.py:57-94 with an all-valid mask -> [h*w, 2*num_feats]
    (row-major over (y, x)); channel order [y-half, x-half]."""
    y = torch.arange(1, h + 1, dtype=torch.float32)[:, None].expand(h, w)
    x = torch.arange(1, w + 1, dtype=torch.float32)[None, :].expand(h, w)
    y = (y / (h + eps) * scale).to(dtype)
    x = (x / (w + eps) * scale).to(dtype)
    return _sine_embed(y.reshape(-1), x.reshape(-1), num_feats, temperature)


def sine_pe_coords(coord, num_feats=128, temperature=10000, scale=2 * math.pi):
    """utils/positional_encoding.py:96-122: coord [B,K,2] (x,y) in [0,1] -> [B,K,2*num_feats]."""
    B, K, _ = coord.shape
    x = (coord[:, :, 0] * scale).reshape(-1)
    y = (coord[:, :, 1] * scale).reshape(-1)
    return _sine_embed(y, x, num_feats, temperature).reshape(B, K, -1)


def _sine_embed(y, x, num_feats, temperature):
    dt = torch.arange(num_feats, dtype=torch.float32)
    dim_t = (temperature ** (2 * torch.div(dt, 2, rounding_mode="floor") / num_feats)).to(y.dtype)
    px = x[:, None] / dim_t
    py = y[:, None] / dim_t
    px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), dim=2).flatten(1)
    py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), dim=2).flatten(1)
    return torch.cat((py, px), dim=1)


def mha(q_in, k_in, v_in, wq, wk, wv, bq, bk, bv, wo, bo, nhead, key_padding_mask=None,
        attn_bias=None):
    """torch.nn.MultiheadAttention arithmetic (batch-first restatement): projections,
    q / sqrt(head_dim), additive -inf key-padding mask, softmax, AV, out_proj.
    attn_bias: optional [B, nhead, Lq, Lk] added before masking (utils/bias_attn.py:183-203)."""
    B, Lq, _ = q_in.shape
    Lk = k_in.shape[1]
    q = F.linear(q_in, wq, bq)
    k = F.linear(k_in, wk, bk)
    v = F.linear(v_in, wv, bv)
    E = q.shape[-1]
    d = E // nhead
    q = q.reshape(B, Lq, nhead, d).transpose(1, 2) * (d ** -0.5)
    k = k.reshape(B, Lk, nhead, d).transpose(1, 2)
    v = v.reshape(B, Lk, nhead, v.shape[-1] // nhead).transpose(1, 2)
    s = q @ k.transpose(-2, -1)
    if attn_bias is not None:
        s = s + attn_bias
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    p = s.softmax(dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Lq, -1)
    return F.linear(o, wo, bo)


def _split3(w):
    n = w.shape[0] // 3
    return w[:n], w[n:2 * n], w[2 * n:]


def gcn(x, adj, w, b):
    """encoder_decoder.py:508-524: Conv1d(k=1) to 2*dff channels, view [B,2,dff,K],
    einsum('bkcv,bkwv->bcw'), ReLU.  x [B,K,d], adj [B,2,K,K] -> [B,K,dff]."""
    B, K, _ = x.shape
    h = F.linear(x, w.reshape(w.shape[0], -1), b)          # [B,K,2*dff]
    dff = h.shape[-1] // 2
    h = h.reshape(B, K, 2, dff)
    y = torch.einsum("bwv,bvc->bwc", adj[:, 0], h[:, :, 0]) + \
        torch.einsum("bwv,bvc->bwc", adj[:, 1], h[:, :, 1])
    return F.relu(y)


def ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def mlp_gelu(x, sd, p, n):
    """encoder_decoder.py:21-34 MLP: Linear+GELU ... Linear."""
    for i in range(n):
        x = F.linear(x, sd[f"{p}.layers.{i}.weight"], sd[f"{p}.layers.{i}.bias"])
        if i < n - 1:
            x = F.gelu(x)
    return x


def kpt_branch(x, sd, p):
    """head.py:34-58 TokenDecodeMLP: 3 x (Linear+GELU) + Linear(->2)."""
    for i in (0, 2, 4):
        x = F.gelu(F.linear(x, sd[f"{p}.mlp.{i}.weight"], sd[f"{p}.mlp.{i}.bias"]))
    return F.linear(x, sd[f"{p}.mlp.6.weight"], sd[f"{p}.mlp.6.bias"])


# ------------------------------------------------------------------- transformer layers
def encoder_layer(src, pos, key_mask, sd, p, nhead):
    """encoder_decoder.py:461-483: pos is added to the residual stream (q, k AND v)."""
    src = src + pos
    wq, wk, wv = _split3(sd[p + ".self_attn.in_proj_weight"])
    bq, bk, bv = _split3(sd[p + ".self_attn.in_proj_bias"])
    a = mha(src, src, src, wq, wk, wv, bq, bk, bv, sd[p + ".self_attn.out_proj.weight"],
            sd[p + ".self_attn.out_proj.bias"], nhead, key_mask)
    src = ln(src + a, sd, p + ".norm1")
    f = F.linear(F.relu(F.linear(src, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                 sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return ln(src + f, sd, p + ".norm2")


def _cross(sd, p, q_in, k_in, v_in, nhead, key_mask=None):
    b = sd[p + ".in_proj_bias"]
    E = sd[p + ".q_proj_weight"].shape[0]
    return mha(q_in, k_in, v_in, sd[p + ".q_proj_weight"], sd[p + ".k_proj_weight"],
               sd[p + ".v_proj_weight"], b[:E], b[E:2 * E], b[2 * E:],
               sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"], nhead, key_mask)


def decoder_layer(kp, img, sd, p, nhead, tgt_key_mask, img_pos, kp_pos, adj, attn_adj=None,
                  two_way=False):
    """encoder_decoder.py:584-651.  kp [B,K,d]; img [B,S,d]; img_pos [S,d] (grid PE);
    kp_pos [B,K,d] = init_pos_emb + concat_pos_embed[S:] (the latter is zero);
    attn_adj [5,B,K,K] or None (biased self-attention, utils/bias_attn.py:106-231)."""
    B, K, d = kp.shape
    if attn_adj is not None:
        hops = attn_adj.permute(1, 2, 3, 0)                                   # [B,K,K,5]
        m = p + ".self_attn.markov_structural_mlp"
        bias = F.linear(F.relu(F.linear(hops, sd[m + ".0.weight"], sd[m + ".0.bias"])),
                        sd[m + ".3.weight"], sd[m + ".3.bias"])               # [B,K,K,H]
        bias = bias.permute(0, 3, 1, 2)
        s = p + ".self_attn"
        a = mha(kp, kp, kp, sd[s + ".q_proj.weight"], sd[s + ".k_proj.weight"],
                sd[s + ".v_proj.weight"], sd[s + ".q_proj.bias"], sd[s + ".k_proj.bias"],
                sd[s + ".v_proj.bias"], sd[s + ".out_proj.weight"], sd[s + ".out_proj.bias"],
                nhead, tgt_key_mask, attn_bias=bias)
    else:
        wq, wk, wv = _split3(sd[p + ".self_attn.in_proj_weight"])
        bq, bk, bv = _split3(sd[p + ".self_attn.in_proj_bias"])
        a = mha(kp, kp, kp, wq, wk, wv, bq, bk, bv, sd[p + ".self_attn.out_proj.weight"],
                sd[p + ".self_attn.out_proj.bias"], nhead, tgt_key_mask)
    kp = ln(kp + a, sd, p + ".norm1")
    pos_b = img_pos[None].expand(B, -1, -1)
    cq = torch.cat((kp, kp_pos), dim=-1)
    ck = torch.cat((img, pos_b), dim=-1)
    a = _cross(sd, p + ".multihead_attn", cq, ck, img, nhead)      # memory mask is all-valid
    kp = ln(kp + F.linear(a, sd[p + ".choker.weight"], sd[p + ".choker.bias"]), sd, p + ".norm2")
    g = gcn(kp, adj, sd[p + ".ffn1.conv.weight"], sd[p + ".ffn1.conv.bias"])
    kp = ln(kp + F.linear(g, sd[p + ".ffn2.weight"], sd[p + ".ffn2.bias"]), sd, p + ".norm3")
    if two_way:
        q2 = torch.cat((img, pos_b), dim=-1)
        k2 = torch.cat((kp, kp_pos), dim=-1)
        a = _cross(sd, p + ".cross_attn_image_to_token", q2, k2, kp, nhead)   # NO key mask (:642-647)
        img = ln(img + F.linear(a, sd[p + ".cross_attn_image_to_token_choker.weight"],
                                sd[p + ".cross_attn_image_to_token_choker.bias"]), sd, p + ".norm4")
    return kp, img


# ------------------------------------------------------------------------- skeleton head
def adj_from_edges(skeleton, K, kp_mask, dtype):
    """keypoint_heads/skeleton.py:171-194: symmetric 0/1 adjacency, masked, row-normalised
    with nan_to_num; stacked with diag(valid)."""
    B = len(skeleton)
    A = torch.zeros(B, K, K, dtype=dtype)
    for b in range(B):
        e = torch.as_tensor(skeleton[b], dtype=torch.long)
        if e.dim() > 1 and e.numel() > 0:
            A[b, e[:, 0], e[:, 1]] = 1
            A[b, e[:, 1], e[:, 0]] = 1
    valid = (~kp_mask).to(dtype)
    A = A * valid[:, :, None] * valid[:, None, :]
    An = torch.nan_to_num(A / A.sum(dim=-1, keepdim=True))
    return torch.stack((torch.diag_embed(valid), An), dim=1)


def soft_normalize_adj(A, kp_mask):
    """skeleton.py:196-205 (mask_res=False, adj_normalization=True, gcn_norm=False)."""
    valid = (~kp_mask).to(A.dtype)
    A = A * (valid[:, :, None] * valid[:, None, :])
    A = A / (A.sum(dim=-1, keepdim=True) + 1e-8)
    return torch.stack((torch.diag_embed(valid), A), dim=1)


def skeleton_forward(sd, p, cfg, skeleton, kp_feat, feats_s, kp_mask, grid_pos, out=None):
    """skeleton.py:58-161.  feats_s: list(shots) of token-major ViT features [B,S,C].
    Returns adj [B,2,K,K], attn_adj [5,B,K,K] or None, unnormalised adj."""
    B, K, d = kp_feat.shape
    dtype = kp_feat.dtype
    gt_adj = adj_from_edges(skeleton, K, kp_mask, dtype)
    binary = gt_adj[:, 1] > 0
    if not cfg.get("learn_skeleton", False):
        return gt_adj, None, binary
    nhead = cfg.get("nhead", 8)
    # refine_features (:82-115) -- uses the binary GT adjacency
    adj_b = soft_normalize_adj(binary.to(dtype), kp_mask)
    m2 = kp_mask.clone()
    m2[(~kp_mask).sum(dim=-1) == 0, 0] = False
    zero_pos = torch.zeros_like(kp_feat)
    w_ip = sd[p + ".image_project.weight"]
    acc = []
    for feat in feats_s:
        img = F.linear(feat, w_ip.reshape(w_ip.shape[0], -1), sd[p + ".image_project.bias"])
        kp = kp_feat
        for i in range(cfg.get("num_layers", 3)):
            kp, img = decoder_layer(kp, img, sd, f"{p}.skeleton_predictor.{i}", nhead, m2, grid_pos,
                                    zero_pos, adj_b, None, two_way=cfg.get("two_way_attn", True))
        acc.append(kp)
    kp = torch.stack(acc).mean(0)
    if out is not None:
        out["skeleton_kp_features"] = kp
    # predict_skeleton (:134-150)
    f = kp / (kp.norm(dim=-1, keepdim=True) + 1e-8)
    S = f @ f.transpose(1, 2)
    S = (S + S.transpose(1, 2)) / 2
    if cfg.get("use_zero_conv", True):
        S = S * sd[p + ".zero_conv.weight"].reshape(()) + sd[p + ".zero_conv.bias"].reshape(())
    U = F.relu(binary.to(dtype) + S)
    adj = soft_normalize_adj(U, kp_mask)
    valid = (~kp_mask).to(dtype)
    U = U * valid[:, :, None] * valid[:, None, :]
    # markov_transition_matrix (:152-161)
    P = adj[:, 1] / (adj[:, 1].sum(dim=-1, keepdim=True) + 1e-8)
    attn_adj = torch.stack([torch.matrix_power(P, h) for h in range(cfg["max_hop"] + 1)])
    return adj, attn_adj, U


# --------------------------------------------------------------------------- proposals
def proposal_generator(sd, p, img, kp, h, w):
    """encoder_decoder.py:49-112.  img [B,S,d], kp [B,K,d] ->
    proposal_for_loss [B,K,2], similarity [B,K,h,w], proposals [B,K,2], argmax [B,K]."""
    B, S, _ = img.shape
    K = kp.shape[1]
    dtype = img.dtype
    fs = F.linear(kp, sd[p + ".support_proj.weight"], sd[p + ".support_proj.bias"])
    fq = F.linear(img, sd[p + ".query_proj.weight"], sd[p + ".query_proj.bias"])
    gate = torch.tanh(F.linear(F.relu(F.linear(fs, sd[p + ".dynamic_proj.0.weight"],
                                               sd[p + ".dynamic_proj.0.bias"])),
                               sd[p + ".dynamic_proj.2.weight"], sd[p + ".dynamic_proj.2.bias"]))
    fsf = (gate + 1) * fs
    sim = torch.bmm(fq, fsf.transpose(1, 2)).transpose(1, 2)             # [B,K,S]
    gy, gx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, dtype=torch.float32),
                            torch.linspace(0.5, w - 0.5, w, dtype=torch.float32), indexing="ij")
    grid = torch.stack((gx, gy), dim=-1).reshape(S, 2).to(dtype)
    norm = torch.tensor([w, h], dtype=dtype)
    sm = sim.softmax(dim=-1)
    prop_loss = (sm[..., None] * grid).sum(dim=2) / norm
    amax = sim.argmax(dim=-1)                                              # first max on ties
    onehot = F.one_hot(amax, S).reshape(B, K, w, h).to(dtype)              # (w,h) quirk (:93)
    local = F.max_pool2d(onehot, kernel_size=3, stride=1, padding=1).reshape(B, K, S)
    lsm = sm * local
    lsm = lsm / (lsm.sum(dim=-1, keepdim=True) + 1e-10)
    prop = (lsm[..., None] * grid).sum(dim=2) / norm
    return prop_loss, sim.reshape(B, K, h, w), prop, amax


# ---------------------------------------------------------------------------------- head
def head_forward(sd, cfg, feat_q, feats_s, target_s, mask_s, skeleton, prefix="keypoint_head_module"):
    """keypoint_heads/head.py:161-222 (+ transformer encoder_decoder.py:183-260, 330-425).
    feat_q / feats_s[i]: token-major ViT features [B,S,C] with S = h*w.  Returns a dict of
    every intermediate SURVEY.md section 8a names."""
    p = prefix
    tcfg = cfg["transformer"]
    nhead = tcfg.get("nhead", 8)
    dtype = feat_q.dtype
    B, S, C = feat_q.shape
    h = w = int(round(math.sqrt(S)))
    assert h * w == S
    out = {}
    w_in = sd[p + ".input_proj.weight"]
    img = F.linear(feat_q, w_in.reshape(w_in.shape[0], -1), sd[p + ".input_proj.bias"])
    d = img.shape[-1]
    grid_pos = sine_pe_grid(h, w, dtype, num_feats=cfg["positional_encoding"]["num_feats"])
    # support keypoint pooling (:175-188)
    pooled = []
    for feat, target in zip(feats_s, target_s):
        hm = target.shape[-1]
        f4 = feat.transpose(1, 2).reshape(B, C, h, w)
        up = F.interpolate(f4, size=target.shape[-2:], mode="bilinear", align_corners=False)
        t = target / (target.sum(dim=-1).sum(dim=-1)[:, :, None, None] + 1e-8)
        pooled.append(t.flatten(2) @ up.flatten(2).permute(0, 2, 1))
    kp = torch.stack(pooled).mean(0) * mask_s
    out["support_keypoints_pooled"] = kp
    kp = F.linear(kp, sd[p + ".query_proj.weight"], sd[p + ".query_proj.bias"])
    out["support_keypoints"] = kp
    kp_mask = ~(mask_s.to(torch.bool).squeeze(-1))
    K = kp.shape[1]
    # skeleton head (:196)
    scfg = dict(cfg["skeleton_head"])
    scfg["max_hop"] = tcfg.get("max_hops", 4)                       # head.py:122
    adj, attn_adj, unnorm = skeleton_forward(sd, p + ".skeleton_head", scfg, skeleton, kp, feats_s,
                                             kp_mask, grid_pos, out)
    out["adj"], out["attn_adj"], out["unnormalized_adj"] = adj, attn_adj, unnorm
    # encoder (encoder_decoder.py:198-203, 276-310)
    t = p + ".transformer"
    x = torch.cat((img, kp), dim=1)
    pos = torch.cat((grid_pos[None].expand(B, -1, -1), torch.zeros(B, K, d, dtype=dtype)), dim=1)
    key_mask = torch.cat((torch.zeros(B, S, dtype=torch.bool), kp_mask), dim=1)
    for i in range(tcfg.get("num_encoder_layers", 3)):
        x = encoder_layer(x, pos, key_mask, sd, f"{t}.encoder.layers.{i}", nhead)
    img, kp = x[:, :S], x[:, S:]
    out["encoder_image"], out["encoder_kp"] = img, kp
    # proposals (:206-210)
    prop_loss, sim, prop, amax = proposal_generator(sd, t + ".proposal_generator", img, kp, h, w)
    out.update(initial_proposals_for_loss=prop_loss, similarity_map=sim, initial_proposals=prop,
               argmax=amax)
    # decoder (:330-425)
    m2 = kp_mask.clone()
    m2[(~kp_mask).sum(dim=-1) == 0, 0] = False
    use_bias = tcfg.get("attn_bias", False)
    bi = prop
    points = [prop]
    hs = []
    L = tcfg.get("num_decoder_layers", 3)
    for i in range(L):
        qpos = mlp_gelu(sine_pe_coords(bi, cfg["positional_encoding"]["num_feats"]), sd,
                        t + ".decoder.ref_point_head", 2)
        kp, img = decoder_layer(kp, img, sd, f"{t}.decoder.layers.{i}", nhead, m2, grid_pos, qpos, adj,
                                attn_adj if use_bias else None, two_way=False)
        hs.append(ln(kp, sd, t + ".decoder.norm"))
        delta = kpt_branch(kp, sd, f"{p}.kpt_branch.{i}")
        bi = (inverse_sigmoid(bi) + delta).sigmoid()
        points.append(bi)
    out["decoder_hs"] = torch.stack(hs)
    out["out_points"] = torch.stack(points)
    # final per-layer decode (head.py:216-220)
    outs = [(kpt_branch(hs[i], sd, f"{p}.kpt_branch.{i}") + inverse_sigmoid(points[i])).sigmoid()
            for i in range(L)]
    out["output"] = torch.stack(outs)
    return out


# ------------------------------------------------------------------------------ detector
def transform_preds(coords, center, scale, output_size):
    """utils/post_processing/post_transforms.py:150-194 (use_udp=False)."""
    sx = scale[0] * 200.0 / output_size[0]
    sy = scale[1] * 200.0 / output_size[1]
    o = coords.clone()
    o[:, 0] = coords[:, 0] * sx + center[0] - scale[0] * 200.0 * 0.5
    o[:, 1] = coords[:, 1] * sy + center[1] - scale[1] * 200.0 * 0.5
    return o


def detector_forward_test(sd, model_cfg, data, dtype=torch.float32, vit_cfg=None):
    """detectors/EdgeCape.py:131-191 + head.decode (head.py:324-387).
    `data` is the forward() kwargs dict (edgecape_b200.synthetic.make_episode)."""
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    vcfg = vit_config(vit_cfg if vit_cfg is not None else model_cfg.get("pretrained", "dinov2_vits14"))
    img_q = data["img_q"].to(dtype)
    img_s = [x.to(dtype) for x in data["img_s"]]
    target_s = [x.to(dtype) for x in data["target_s"]]
    mask_s = data["target_weight_s"][0].to(dtype)
    for tw in data["target_weight_s"]:
        mask_s = mask_s * tw.to(dtype)                          # EdgeCape.py:175-177 (first twice)
    feat_q, _ = vit_forward_tokens(sd, vcfg, img_q, "encoder_query.")
    feats_s = [vit_forward_tokens(sd, vcfg, x, "encoder_sample.")[0] for x in img_s]
    skeleton = [m["sample_skeleton"][0] for m in data["img_metas"]]
    out = head_forward(sd, model_cfg["keypoint_head"], feat_q, feats_s, target_s, mask_s, skeleton)
    out["feature_q"] = feat_q
    B = img_q.shape[0]
    W, H = img_q.shape[-1], img_q.shape[-2]
    pose = out["output"][-1] * torch.tensor([W, H], dtype=dtype)
    preds = torch.zeros(B, pose.shape[1], 3, dtype=dtype)
    boxes = torch.zeros(B, 6, dtype=dtype)
    for i, m in enumerate(data["img_metas"]):
        c = torch.as_tensor(m["query_center"], dtype=dtype)
        s = torch.as_tensor(m["query_scale"], dtype=dtype)
        preds[i, :, :2] = transform_preds(pose[i], c, s, [W, H])
        preds[i, :, 2] = 1.0
        boxes[i, 0:2], boxes[i, 2:4] = c, s
        boxes[i, 4] = torch.prod(s * 200.0)
        boxes[i, 5] = float(m.get("query_bbox_score", 1.0))
    out["preds"], out["boxes"] = preds, boxes
    out["points"] = torch.cat((out["initial_proposals_for_loss"][None], out["output"]))
    out["skeleton"] = out["adj"][0]
    return out
