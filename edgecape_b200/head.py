"""@HEADS TwoStageHead on the edgecape_b200 kernels.

Same constructor kwargs, `forward` / `decode` signatures and state-dict keys as
/root/reference/EdgeCape/models/keypoint_heads/head.py:61-387 (inference part: `forward` :161-222,
`decode` :324-387; the losses :224-322 are training-only and out of scope).
"""
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .config import _lin_keys
from .params import PackedMixin, ParamTree, xavier_uniform_all_
from .registry import HEADS, build_head, build_positional_encoding, build_transformer
from .transformer import token_decode_mlp


def transform_preds(coords, center, scale, output_size, use_udp=False):
    """mmpose.core.post_processing.transform_preds (spec: the vendored copy at
    /root/reference/EdgeCape/models/utils/post_processing/post_transforms.py:150-194)."""
    assert coords.shape[1] in (2, 4, 5)
    assert len(center) == 2 and len(scale) == 2 and len(output_size) == 2
    scale = scale * 200.0
    if use_udp:
        scale_x = scale[0] / (output_size[0] - 1.0)
        scale_y = scale[1] / (output_size[1] - 1.0)
    else:
        scale_x = scale[0] / output_size[0]
        scale_y = scale[1] / output_size[1]
    target = np.ones_like(coords)
    target[:, 0] = coords[:, 0] * scale_x + center[0] - scale[0] * 0.5
    target[:, 1] = coords[:, 1] * scale_y + center[1] - scale[1] * 0.5
    return target


@HEADS.register_module(force=True)
class TwoStageHead(PackedMixin, nn.Module):
    def __init__(self, in_channels, transformer=None,
                 positional_encoding=dict(type="SinePositionalEncoding", num_feats=128, normalize=True),
                 share_kpt_branch=False, num_decoder_layer=3, with_heatmap_loss=False, heatmap_loss_weight=2.0,
                 skeleton_loss_weight=1, train_cfg=None, test_cfg=None, skeleton_head=None, learn_skeleton=False,
                 masked_supervision=False, freeze=None, model_freeze=None, masking_ratio=0.5):
        super().__init__()
        self._init_packed()
        self.in_channels = in_channels
        self.positional_encoding = build_positional_encoding(positional_encoding)
        self.encoder_positional_encoding = build_positional_encoding(positional_encoding)
        self.transformer = build_transformer(transformer)
        self.embed_dims = self.transformer.d_model
        self.with_heatmap_loss = with_heatmap_loss
        self.heatmap_loss_weight = heatmap_loss_weight
        self.skeleton_loss_weight = skeleton_loss_weight
        assert "num_feats" in positional_encoding
        num_feats = positional_encoding["num_feats"]
        assert num_feats * 2 == self.embed_dims, \
            f"embed_dims should be exactly 2 times of num_feats. Found {self.embed_dims} and {num_feats}."
        d = self.embed_dims
        top = {"input_proj.weight": (d, in_channels, 1, 1), "input_proj.bias": (d,)}
        top.update(_lin_keys("query_proj", d, in_channels))
        tree = ParamTree(top)
        self.input_proj, self.query_proj = tree.input_proj, tree.query_proj
        kb = {}
        for j in (0, 2, 4):
            kb.update(_lin_keys(f"mlp.{j}", d, d))
        kb.update(_lin_keys("mlp.6", 2, d))
        branch = ParamTree(kb)
        self.share_kpt_branch = share_kpt_branch
        self.kpt_branch = nn.ModuleList([branch if share_kpt_branch else deepcopy(branch)
                                         for _ in range(num_decoder_layer)])
        self.train_cfg = {} if train_cfg is None else train_cfg
        self.test_cfg = {} if test_cfg is None else test_cfg
        self.target_type = self.test_cfg.get("target_type", "GaussianHeatMap")
        skeleton_head = dict(skeleton_head)
        skeleton_head["max_hop"] = transformer.get("max_hops", 4)                # head.py:122
        self.skeleton_head = build_head(skeleton_head)
        self.learn_skeleton = learn_skeleton
        self.masking_ratio = masking_ratio
        self.masked_supervision = masked_supervision
        self.transformer.masked_supervision = masked_supervision
        self.transformer.mask_token = nn.Parameter(torch.zeros(1, transformer.get("d_model", 256)),
                                                   requires_grad=False)        # head.py:130
        self.transformer.masking_ratio = masking_ratio
        self.use_zero_conv = skeleton_head.get("use_zero_conv", False)
        self.freeze, self.model_freeze = freeze, model_freeze
        self.concurrent_skeleton = bool(self.test_cfg.get("concurrent_skeleton", True))
        self._side = None

    def init_weights(self):
        """head.py:143-159 (xavier everywhere, zero last kpt_branch layer and zero_conv)."""
        xavier_uniform_all_(self)
        with torch.no_grad():
            for mlp in self.kpt_branch:
                getattr(mlp.mlp, "6").weight.zero_()
                getattr(mlp.mlp, "6").bias.zero_()
            self.input_proj.bias.zero_()
            self.query_proj.bias.zero_()
            if self.use_zero_conv and hasattr(self.skeleton_head, "zero_conv"):
                self.skeleton_head.zero_conv.weight.zero_()
                self.skeleton_head.zero_conv.bias.zero_()
            for n, p in self.named_parameters():
                if "norm" in n and n.endswith("weight"):
                    p.fill_(1.0)
                elif n.endswith("bias") and "norm" in n:
                    p.zero_()
        self.invalidate_packed()

    # ---------------------------------------------------------------------------- forward
    @torch.no_grad()
    def forward_tokens(self, feat_q, feats_s, target_s, mask_s, skeleton_lst, return_intermediates=False):
        """Token-major entry used by the detector.

        feat_q [B,S,C], feats_s list(shots) of [B,S,C] (views into the ViT token buffer are fine),
        target_s list(shots) of [B,K,hm,hm], mask_s [B,K] float (product of the visibility weights),
        skeleton_lst list(batch) of edge lists.  Returns the reference's 5-tuple
        (output [L,B,K,2], initial_proposals_for_loss [B,K,2], similarity_map [B,K,h,w], None,
        adj [B,2,K,K]) plus, on request, a dict of intermediates named as in SURVEY.md section 8a."""
        B, S, C = feat_q.shape
        h = w = int(round(S ** 0.5))
        assert h * w == S, "square feature maps only (the reference's ProposalGenerator assumes h == w)"
        K = target_s[0].shape[1]
        d = self.embed_dims
        dev = feat_q.device
        shots = len(feats_s)
        T = S + K
        grid_pos = self.positional_encoding.grid_tokens(h, w, dev)                    # [S,d]
        kp_mask, kp_mask_fixed = ops.kp_masks(mask_s)
        # x = [input_proj(feat_q) ; support keypoint tokens]  (head.py:169, encoder_decoder.py:198-203)
        x = ops.empty(B, T, d, device=dev)
        ops.linear(feat_q, self.input_proj.weight, self.input_proj.bias, out=x[:, :S, :])
        # support keypoint pooling (head.py:175-187), exact by linearity of the bilinear resize
        rowscale = ops.axpby(mask_s, mask_s, 1.0, 0.0, float(shots)) if shots > 1 else mask_s
        pooled = None
        for feat, target in zip(feats_s, target_s):
            tw = ops.support_weights(target.contiguous(), rowscale, h, w)               # [B,K,S]
            pooled = ops.gemm(tw, feat, b_kmajor=False, residual=pooled, res_mode=ops.RES_ADD)   # [B,K,C]
        ops.linear(pooled, self.query_proj.weight, self.query_proj.bias, out=x[:, S:, :])
        kp_tokens = ops.copy_rows(x[:, S:, :], ops.empty(B, K, d, device=dev))
        # skeleton / edge-weight predictor (head.py:196) on a side stream: it only depends on the support
        # features and the pooled keypoint tokens, so it overlaps the query encoder + proposal generator; the
        # decoder (first consumer of adj / attn_adj) joins.  Inside a CUDA graph this becomes a parallel branch.
        main = torch.cuda.current_stream(dev) if dev.type == "cuda" else None
        sk = {}
        if main is not None and self.concurrent_skeleton:
            if self._side is None or self._side.device != dev or self._side.priority != main.priority:
                self._side = torch.cuda.Stream(device=dev, priority=main.priority)   # same priority as its parent
            side = self._side
            side.wait_stream(main)
            with torch.cuda.stream(side):
                sk["out"] = self.skeleton_head.forward_tokens(skeleton_lst, kp_tokens, feats_s, kp_mask, kp_mask_fixed,
                                                              grid_pos)

            def join():
                main.wait_stream(side)
                if not torch.cuda.is_current_stream_capturing():   # (a graph's private pool needs no stream tracking)
                    for t_ in sk["out"]:
                        if torch.is_tensor(t_):
                            t_.record_stream(main)
                return sk["out"][0], sk["out"][1]
        else:
            sk["out"] = self.skeleton_head.forward_tokens(skeleton_lst, kp_tokens, feats_s, kp_mask, kp_mask_fixed,
                                                          grid_pos)

            def join():
                return sk["out"][0], sk["out"][1]
        # encoder -> proposals -> graph decoder (head.py:203)
        tr = self.transformer.forward_tokens(x, S, (h, w), grid_pos, kp_mask, kp_mask_fixed,
                                             self.positional_encoding, self.kpt_branch, join, None)
        adj, attn_adj, unnorm, refined = sk["out"]
        # final per-layer decode (head.py:216-220)
        L = tr["hs"].shape[0]
        output = ops.empty(L, B, K, 2, device=dev)
        for i in range(L):
            hs_i = tr["hs_split"][i] if tr["hs_split"][i] is not None else tr["hs"][i].view(B * K, d)
            delta = token_decode_mlp(hs_i, self.kpt_branch[i])
            ops.point_update(tr["out_points"][i], delta, out=output[i])
        res = (output, tr["proposal_for_loss"], tr["similarity_map"], None, adj)
        if not return_intermediates:
            return res
        inter = dict(support_keypoints=kp_tokens, support_keypoints_pooled=pooled, skeleton_kp_features=refined,
                     adj=adj, attn_adj=attn_adj, unnormalized_adj=unnorm, encoder_image=tr["encoder_image"],
                     encoder_kp=tr["encoder_kp"], initial_proposals_for_loss=tr["proposal_for_loss"],
                     similarity_map=tr["similarity_map"], initial_proposals=tr["proposals"], argmax=tr["argmax"],
                     decoder_hs=tr["hs"], out_points=torch.stack(tr["out_points"]), output=output, kp_mask=kp_mask)
        return res, inter

    def forward(self, feature_q, feature_s, target_s, mask_s, skeleton_lst, return_attn_maps=False,
                random_mask=None):
        """Reference signature (head.py:161-168): NCHW feature maps; mask_s [B,K,1]."""
        B, C, h, w = feature_q.shape
        fq = feature_q.flatten(2).transpose(1, 2).contiguous()
        fs = [f.flatten(2).transpose(1, 2).contiguous() for f in feature_s]
        return self.forward_tokens(fq, fs, list(target_s), mask_s.reshape(B, -1).contiguous().float(), skeleton_lst)

    # ----------------------------------------------------------------------------- decode
    def assemble_result(self, img_metas, preds):
        """The bookkeeping half of `decode` (head.py:341-387) around device-decoded `preds` [B,K,3]."""
        batch_size = len(img_metas)
        bbox_ids = []
        c = np.zeros((batch_size, 2), dtype=np.float32)
        s = np.zeros((batch_size, 2), dtype=np.float32)
        image_paths = []
        score = np.ones(batch_size)
        for i in range(batch_size):
            c[i, :] = img_metas[i]["query_center"]
            s[i, :] = img_metas[i]["query_scale"]
            image_paths.append(img_metas[i]["query_image_file"])
            if "query_bbox_score" in img_metas[i]:
                score[i] = np.array(img_metas[i]["query_bbox_score"]).reshape(-1)[0]
            if "bbox_id" in img_metas[i]:
                bbox_ids.append(img_metas[i]["bbox_id"])
            elif "query_bbox_id" in img_metas[i]:
                bbox_ids.append(img_metas[i]["query_bbox_id"])
        all_boxes = np.zeros((batch_size, 6), dtype=np.float32)
        all_boxes[:, 0:2] = c[:, 0:2]
        all_boxes[:, 2:4] = s[:, 0:2]
        all_boxes[:, 4] = np.prod(s * 200.0, axis=1)
        all_boxes[:, 5] = score
        return dict(preds=np.ascontiguousarray(preds, dtype=np.float32), boxes=all_boxes, image_paths=image_paths,
                    bbox_ids=bbox_ids)

    def decode(self, img_metas, output, img_size, **kwargs):
        """head.py:324-387 for host arrays (the default path decodes on the device, ec_decode_preds): scale the
        normalised coordinates to pixels, undo the top-down crop for the whole batch at once (transform_preds
        broadcast over the rows), then the same bookkeeping as assemble_result."""
        W, H = img_size
        px = np.asarray(output, dtype=np.float64) * np.array([W, H])[None, None, :]
        c = np.stack([np.asarray(m["query_center"], dtype=np.float32).reshape(-1)[:2] for m in img_metas])
        s = np.stack([np.asarray(m["query_scale"], dtype=np.float32).reshape(-1)[:2] for m in img_metas]) * 200.0
        denom = np.array([W, H], dtype=np.float64) - (1.0 if self.test_cfg.get("use_udp", False) else 0.0)
        preds = np.ones(px.shape[:2] + (3,), dtype=np.float32)
        preds[:, :, 0:2] = px[:, :, 0:2] * (s / denom)[:, None, :] + (c - 0.5 * s)[:, None, :]
        return self.assemble_result(img_metas, preds)
