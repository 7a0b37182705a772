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
from .skeleton import edges_to_csr, edges_to_csr_host
from .vit import DinoVisionTransformerB200


class _GraphedForward:
    """Captured CUDA graphs of EdgeCape._forward_device for a fixed input signature, with static input
    buffers (images, heat-maps, visibility weights, CSR edge lists up to `edge_capacity`).

    Two graphs: (A) the batched ViT over [query; supports], (B) mask + head.  Per call the images are
    copied on the launch stream and A is replayed, while heat-maps / weights / edge lists (half of the
    H2D bytes, only needed by B) are copied on a side stream concurrently with A; B joins them."""

    def __init__(self, model, img_q, img_s, target_s, target_weight_s, e_np, o_np, dev):
        mk = lambda t: torch.empty(tuple(t.shape), dtype=torch.float32, device=dev)
        self.img_q = mk(img_q)
        self.img_s = [mk(t) for t in img_s]
        self.target_s = [mk(t) for t in target_s]
        self.tw_s = [mk(t) for t in target_weight_s]
        B = img_q.shape[0]
        self.B = B
        self.edge_capacity = max(1024, 2 * int(e_np.shape[0]))
        self.edges = torch.zeros(self.edge_capacity, 2, dtype=torch.int32, device=dev)
        self.offsets = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        self.host = torch.empty(2 * self.edge_capacity + B + 1, dtype=torch.int32).pin_memory()
        self.staged = None
        self.copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        self._load_images(img_q, img_s)
        self._load_head_inputs(main, target_s, target_weight_s, e_np, o_np)
        main.wait_stream(self.copy_stream)
        # warm-up on a side stream (packs weights, fills caches, sets kernel attributes), then capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for _ in range(2):
                model._forward_device(self.img_q, self.img_s, self.target_s, self.tw_s, (self.edges, self.offsets))
        main.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph_vit = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_vit):
            self.feat_q, self.feats_s = model.extract_features(self.img_s, self.img_q)
        self.graph_head = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_head, pool=self.graph_vit.pool()):
            self.out = model._head_device(self.feat_q, self.feats_s, self.target_s, self.tw_s,
                                          (self.edges, self.offsets))

    def _load_images(self, img_q, img_s):
        self.img_q.copy_(img_q, non_blocking=True)
        for d, s_ in zip(self.img_s, img_s):
            d.copy_(s_, non_blocking=True)

    def _load_head_inputs(self, main, target_s, target_weight_s, e_np, o_np):
        if self.staged is not None:
            self.staged.synchronize()      # the pinned CSR staging buffer is reused: wait for its last H2D
        ne, no = e_np.shape[0], o_np.shape[0]
        self.host[:no] = torch.from_numpy(o_np)
        self.host[no:no + 2 * ne] = torch.from_numpy(e_np.reshape(-1))
        cs = self.copy_stream
        cs.wait_stream(main)               # the previous replay of graph B may still read these buffers
        with torch.cuda.stream(cs):
            for d, s_ in zip(self.target_s + self.tw_s, list(target_s) + list(target_weight_s)):
                d.copy_(s_, non_blocking=True)
            self.offsets.copy_(self.host[:no], non_blocking=True)
            if ne:
                self.edges[:ne].view(-1).copy_(self.host[no:no + 2 * ne], non_blocking=True)
            self.staged = torch.cuda.Event()
            self.staged.record(cs)

    def run(self, img_q, img_s, target_s, target_weight_s, e_np, o_np):
        main = torch.cuda.current_stream(self.img_q.device)
        # images first: the DMA engine serves copies in issue order and graph A only needs the images; the
        # head inputs follow on the copy stream and overlap graph A
        self._load_images(img_q, img_s)
        self._load_head_inputs(main, target_s, target_weight_s, e_np, o_np)
        self.graph_vit.replay()
        main.wait_stream(self.copy_stream)
        self.graph_head.replay()
        return self.out


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
        self.use_cuda_graph = bool(self.test_cfg.get("cuda_graph", True))
        self._graphs = {}
        self._host_out = {}
        self.eval()

    def _apply(self, fn, *a, **k):
        self._graphs = {}           # captured graphs hold the old parameter addresses
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._graphs = {}           # repacked weight copies (split fp16, fused QKV) change
        return super().load_state_dict(*a, **k)

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
        # decode on the device (heat-map space -> image space) and ONE synchronisation for every device->host read
        # (the reference does three .cpu() calls and decodes on the host)
        L, B, K, _ = output.shape
        dev = output.device
        host = self._host_out.get((L, B, K))
        if host is None:
            host = (torch.empty(L + 1, B, K, 2), torch.empty(2, K, K), torch.empty(B, K, 3), torch.empty(B, 4))
            if output.is_cuda:
                host = tuple(t.pin_memory() for t in host)
            self._host_out = {(L, B, K): host}
        cs = host[3]
        for i in range(B):
            cs[i, 0:2] = torch.as_tensor(np.asarray(img_metas[i]["query_center"], dtype=np.float32).reshape(-1)[:2])
            cs[i, 2:4] = torch.as_tensor(np.asarray(img_metas[i]["query_scale"], dtype=np.float32).reshape(-1)[:2])
        preds_dev = ops.decode_preds(output[-1].contiguous(), cs.to(dev, non_blocking=True), img_width, img_height,
                                     use_udp=self.keypoint_head_module.test_cfg.get("use_udp", False))
        host[0][0].copy_(initial_proposals, non_blocking=True)
        host[0][1:].copy_(output, non_blocking=True)
        host[1].copy_(adj[0], non_blocking=True)
        host[2].copy_(preds_dev, non_blocking=True)
        if output.is_cuda:
            torch.cuda.current_stream(dev).synchronize()
        points = host[0].numpy().copy()
        result = {}
        if self.with_keypoint:
            result.update(self.keypoint_head_module.assemble_result(img_metas, host[2].numpy().copy()))
        result.update({"points": points})
        result.update({"sample_image_file": [img_metas[i]["sample_image_file"] for i in range(len(img_metas))]})
        result.update({"skeleton": host[1].numpy().copy()})
        return result

    @torch.no_grad()
    def predict(self, img_s, target_s, target_weight_s, img_q, img_metas=None, return_intermediates=False):
        """(:165-184).  Accepts CPU or CUDA tensors; CPU inputs are uploaded (pinned or pageable).
        With `use_cuda_graph` (default; test_cfg['cuda_graph']=False disables) the whole device-side
        forward of one input signature is captured once into a CUDA graph and replayed, which removes
        the ~300 per-launch host costs of a step; intermediates are only available eagerly."""
        dev = self.device
        _require_cuda(dev)
        skeleton_lst = [i["sample_skeleton"][0] for i in img_metas]
        if self.use_cuda_graph and not return_intermediates and dev.type == "cuda":
            return self._predict_graphed(img_s, target_s, target_weight_s, img_q, skeleton_lst)
        up = lambda t: t.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        edges, offsets = edges_to_csr(skeleton_lst, dev)
        return self._forward_device(up(img_q), [up(t) for t in img_s], [up(t) for t in target_s],
                                    [up(t) for t in target_weight_s], (edges, offsets), return_intermediates)

    def _forward_device(self, img_q, img_s, target_s, target_weight_s, skeleton, return_intermediates=False):
        """Device-resident forward: every argument is a CUDA tensor, `skeleton` = (edges, offsets) CSR."""
        dev = img_q.device
        B, K = target_weight_s[0].shape[:2]
        # mask_s = prod of visibility weights; the first one enters twice (:175-177)
        mask_s = ops.empty(B, K, device=dev)
        ops.mask_accumulate_(target_weight_s[0].reshape(B, K), mask_s, first=True)
        for tw in target_weight_s[1:]:
            ops.mask_accumulate_(tw.reshape(B, K), mask_s, first=False)
        feat_q, feats_s = self.extract_features(img_s, img_q)
        return self.keypoint_head_module.forward_tokens(feat_q, feats_s, target_s, mask_s, skeleton,
                                                        return_intermediates=return_intermediates)

    def _head_device(self, feat_q, feats_s, target_s, target_weight_s, skeleton):
        """The part of _forward_device after the backbone (captured as its own CUDA graph)."""
        dev = feat_q.device
        B, K = target_weight_s[0].shape[:2]
        mask_s = ops.empty(B, K, device=dev)
        ops.mask_accumulate_(target_weight_s[0].reshape(B, K), mask_s, first=True)
        for tw in target_weight_s[1:]:
            ops.mask_accumulate_(tw.reshape(B, K), mask_s, first=False)
        return self.keypoint_head_module.forward_tokens(feat_q, feats_s, target_s, mask_s, skeleton)

    # ------------------------------------------------------------------------ CUDA graphs
    def _predict_graphed(self, img_s, target_s, target_weight_s, img_q, skeleton_lst):
        dev = self.device
        e_np, o_np = edges_to_csr_host(skeleton_lst)
        key = (tuple(img_q.shape), len(img_s), tuple(target_s[0].shape), ops.TENSOR_CORES)
        g = self._graphs.get(key)
        if g is None or g.edge_capacity < e_np.shape[0]:
            g = _GraphedForward(self, img_q, img_s, target_s, target_weight_s, e_np, o_np, dev)
            if len(self._graphs) >= 8:
                self._graphs.clear()
            self._graphs[key] = g
        return g.run(img_q, img_s, target_s, target_weight_s, e_np, o_np)

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
