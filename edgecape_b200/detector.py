"""@POSENETS EdgeCape detector -- drop-in for
/root/reference/EdgeCape/models/detectors/EdgeCape.py:17-191 (inference: `forward` :56-80,
`forward_test` :131-163, `predict` :165-184, `extract_features` :186-191).

The backbone is this package's DINOv2 implementation instead of `torch.hub.load` (no network, no
hub code); it is bound to both `encoder_sample` and `encoder_query` exactly like the reference
(one module, two state-dict prefixes, :36).  Query and support images go through the ViT as ONE
batch, and every op runs in the CUDA library.  Training (`forward_train`, :82-129) is out of scope.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from .registry import POSENETS, build_head
from .vit import DinoVisionTransformerB200


def _require_cuda(dev):
    if dev.type != "cuda":
        raise _lib.EdgeCapeLibraryError(
            "EdgeCape (edgecape_b200) runs on CUDA only: move the model with .cuda() first; there is no CPU path")


@POSENETS.register_module(force=True)
class EdgeCape(nn.Module):
    def __init__(self, keypoint_head, encoder_config, train_cfg=None, test_cfg=None, pretrained="dinov2_vits14"):
        super().__init__()
        self.encoder_sample = self.encoder_query = DinoVisionTransformerB200(pretrained)
        self.backbone = "dinov2"
        self.keypoint_head_module = build_head(keypoint_head)
        self.keypoint_head_module.init_weights()
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg if test_cfg is not None else {}
        self.target_type = self.test_cfg.get("target_type", "GaussianHeatMap")
        self.eval()

    @property
    def with_keypoint(self):
        return hasattr(self, "keypoint_head_module")

    def init_weights(self):
        self.keypoint_head_module.init_weights()

    @property
    def device(self):
        return next(self.parameters()).device

    # ----------------------------------------------------------------------------- API
    def forward(self, img_s, img_q, target_s=None, target_weight_s=None, target_q=None, target_weight_q=None,
                img_metas=None, return_loss=True, **kwargs):
        """Same dispatch as the reference (:56-80); only the test branch exists here."""
        if return_loss:
            raise NotImplementedError(
                "edgecape_b200 accelerates the inference path only: call model(return_loss=False, **data) "
                "(forward_train / losses are out of scope, see DESIGN.md)")
        return self.forward_test(img_s, target_s, target_weight_s, img_q, target_q, target_weight_q, img_metas,
                                 **kwargs)

    @torch.no_grad()
    def forward_test(self, img_s, target_s, target_weight_s, img_q, target_q, target_weight_q, img_metas=None,
                     vis_offset=True, **kwargs):
        """Returns the reference's result dict (:131-163): preds [B,K,3], boxes [B,6], image_paths,
        bbox_ids, points [1+L,B,K,2], sample_image_file, skeleton [2,K,K] (adjacency of batch item 0)."""
        batch_size, _, img_height, img_width = img_q.shape
        output, initial_proposals, similarity_map, _, adj = self.predict(img_s, target_s, target_weight_s, img_q,
                                                                         img_metas)
        predicted_pose = output[-1].detach().cpu().numpy()
        result = {}
        if self.with_keypoint:
            keypoint_result = self.keypoint_head_module.decode(img_metas, predicted_pose,
                                                               img_size=[img_width, img_height])
            result.update(keypoint_result)
        result.update({"points": torch.cat((initial_proposals[None], output)).cpu().numpy()})
        result.update({"sample_image_file": [img_metas[i]["sample_image_file"] for i in range(len(img_metas))]})
        result.update({"skeleton": adj[0].cpu().numpy()})
        return result

    @torch.no_grad()
    def predict(self, img_s, target_s, target_weight_s, img_q, img_metas=None, return_intermediates=False):
        """(:165-184).  Accepts CPU or CUDA tensors; CPU inputs are uploaded (pinned or pageable)."""
        dev = self.device
        _require_cuda(dev)
        up =lambda t: t.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        img_q = up(img_q)
        img_s = [up(t) for t in img_s]
        target_s = [up(t) for t in target_s]
        target_weight_s = [up(t) for t in target_weight_s]
        B, K = target_weight_s[0].shape[:2]
        # mask_s = prod of visibility weights; the first one enters twice (:175-177)
        mask_s = ops.empty(B, K, device=dev)
        ops.mask_accumulate_(target_weight_s[0].reshape(B, K), mask_s, first=True)
        for tw in target_weight_s[1:]:
            ops.mask_accumulate_(tw.reshape(B, K), mask_s, first=False)
        feat_q, feats_s = self.extract_features(img_s, img_q)
        skeleton_lst = [i["sample_skeleton"][0] for i in img_metas]
        return self.keypoint_head_module.forward_tokens(feat_q, feats_s, target_s, mask_s, skeleton_lst,
                                                        return_intermediates=return_intermediates)

    @torch.no_grad()
    def extract_features(self, img_s, img_q):
        """(:186-191) one batched ViT pass over [query; support shots]; returns token-major views
        feat_q [B,S,C] and a list of feats_s [B,S,C] (cls row dropped by striding, no copy)."""
        tok, _ = self.encoder_query.forward_tokens([img_q] + list(img_s))
        B = img_q.shape[0]
        feat_q = tok[:B, 1:, :]
        feats_s = [tok[(i + 1) * B:(i + 2) * B, 1:, :] for i in range(len(img_s))]
        return feat_q, feats_s

    def forward_train(self, *args, **kwargs):
        raise NotImplementedError("training is out of scope for edgecape_b200 (inference hot path only)")

    def show_result(self, *args, **kwargs):
        raise NotImplementedError("visualisation is out of scope for edgecape_b200")
