"""Deterministic synthetic weights and episodes.

No dataset or checkpoint is reachable offline, so parity tests, goldens, smoke() and
bench.py all draw weights and inputs from the seeded generators below.  Everything
is produced with numpy's PCG64 (`default_rng`), which is stable across machines and
numpy versions, then wrapped as torch CPU tensors.

* `make_state_dict(shapes, seed)` fills a {key: shape} map.  Matrices get
  U(-a, a) with a = sqrt(6/(fan_in+fan_out)) (xavier-uniform scale, the reference's
  own init, head.py:143-159) -- ViT matrices N(0, 0.02) (upstream trunc-normal scale);
  LayerNorm weights 1+0.1 N, biases 0.02 N, LayerScale gamma 1+0.1 N.  The reference
  zero-initialises `zero_conv` and the last `kpt_branch` layers (head.py:151-159),
  which would hide the edge-weight and refinement paths, so they are randomised too.
* `make_episode(...)` builds one batch of the forward() data dict
  (/root/reference/demo.py:205-228 is the minimal contract): N(0,1) images, uniform
  keypoints, 64x64 sigma=1 MSRA Gaussian support heat-maps
  (datasets/pipelines/top_down_transform.py:165-194 semantics), visibility weights and a
  random skeleton (spanning tree + extra edges, or fully connected).
"""
import zlib

import numpy as np
import torch


def _rng(seed, key):
    return np.random.default_rng([int(seed), zlib.crc32(key.encode())])


def make_tensor(key, shape, seed):
    """One deterministic fp32 tensor for state-dict entry `key`."""
    r = _rng(seed, key)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    is_vit = key.startswith(("encoder_query.", "encoder_sample.")) or key.startswith("vit.")
    if "zero_conv" in key:
        a = np.full(shape, 0.6 if leaf == "weight" else 0.05, dtype=np.float32)
    elif leaf == "gamma":
        a = 1.0 + 0.1 * r.standard_normal(shape)
    elif "norm" in key and leaf == "weight" and len(shape) == 1:
        a = 1.0 + 0.1 * r.standard_normal(shape)
    elif len(shape) <= 1 or leaf == "bias":
        a = 0.02 * r.standard_normal(shape)
    elif leaf in ("cls_token", "pos_embed", "mask_token"):
        a = 0.02 * r.standard_normal(shape)
    elif is_vit:
        a = 0.02 * r.standard_normal(shape)
    else:
        fan_out = shape[0]
        fan_in = int(np.prod(shape[1:]))
        bound = np.sqrt(6.0 / (fan_in + fan_out))
        if "kpt_branch" in key and key.endswith("mlp.6.weight"):
            bound *= 0.25       # keeps sigmoid updates away from saturation
        a = r.uniform(-bound, bound, size=shape)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def make_state_dict(shapes, seed=0):
    """shapes: {key: shape}.  `encoder_sample.*` aliases `encoder_query.*` (one backbone,
    two names -- detectors/EdgeCape.py:36)."""
    sd = {}
    for k in sorted(shapes):
        kk = k.replace("encoder_sample.", "encoder_query.")
        sd[k] = make_tensor(kk, shapes[k], seed)
    return sd


def msra_target(xy, image_size, heatmap_size=64, sigma=1.0):
    """xy [K,2] pixel coords -> un-normalised Gaussian targets [K,hm,hm] (MSRA, biased
    encoding: integer centre, (6 sigma + 1)^2 patch clipped at the border)."""
    K = xy.shape[0]
    W = H = heatmap_size
    target = np.zeros((K, H, W), dtype=np.float32)
    tmp = int(sigma * 3)
    size = 2 * tmp + 1
    x = np.arange(0, size, 1, np.float32)
    y = x[:, None]
    g = np.exp(-((x - size // 2) ** 2 + (y - size // 2) ** 2) / (2 * sigma ** 2)).astype(np.float32)
    stride = image_size / heatmap_size
    for j in range(K):
        mu_x = int(xy[j, 0] / stride + 0.5)
        mu_y = int(xy[j, 1] / stride + 0.5)
        ul = [mu_x - tmp, mu_y - tmp]
        br = [mu_x + tmp + 1, mu_y + tmp + 1]
        if ul[0] >= W or ul[1] >= H or br[0] < 0 or br[1] < 0:
            continue
        gx = max(0, -ul[0]), min(br[0], W) - ul[0]
        gy = max(0, -ul[1]), min(br[1], H) - ul[1]
        ix = max(0, ul[0]), min(br[0], W)
        iy = max(0, ul[1]), min(br[1], H)
        target[j, iy[0]:iy[1], ix[0]:ix[1]] = g[gy[0]:gy[1], gx[0]:gx[1]]
    return target


def random_skeleton(rng, K, n_valid=None, kind="tree+extra"):
    """Edge list (0-based index pairs) over the first `n_valid` of K keypoints."""
    n = K if n_valid is None else n_valid
    if kind == "chain":
        return [[i, i + 1] for i in range(n - 1)]
    if kind == "full":
        return [[i, j] for i in range(n) for j in range(i + 1, n)]
    if kind == "empty":
        return []
    edges = []
    perm = rng.permutation(n)
    for i in range(1, n):
        edges.append([int(perm[i]), int(perm[rng.integers(0, i)])])
    for _ in range(n // 2):
        a, b = rng.integers(0, n, size=2)
        if a != b:
            edges.append([int(a), int(b)])
    return edges


def make_episode(batch, image_size=256, num_kpts=100, shots=1, seed=1234, masked_tail=0.0,
                 skeleton="tree+extra", heatmap_size=64, pin_memory=False, shared_support=1):
    """One batch of the reference's forward() kwargs (CPU tensors).

    masked_tail: fraction of trailing keypoints marked invisible (MP-100 pads categories to
    100 keypoints, datasets/datasets/mp100/test_dataset.py:187-197).
    shared_support: runs of this many consecutive rows share their support sample -- image(s), heat-maps, skeleton and
    `sample_image_file` -- as the queries of one MP-100 test episode do (test_dataset.py:93-97: 15 queries per support)."""
    rng = np.random.default_rng([int(seed), 77])
    B, K, R = batch, num_kpts, image_size
    n_valid = K - int(round(masked_tail * K))
    img_q = rng.standard_normal((B, 3, R, R), dtype=np.float32)
    img_s = [rng.standard_normal((B, 3, R, R), dtype=np.float32) for _ in range(shots)]
    target_s, weight_s, kpts_s = [], [], []
    for _ in range(shots):
        xy = rng.uniform(0.05 * R, 0.95 * R, size=(B, K, 2)).astype(np.float32)
        t = np.stack([msra_target(xy[b], R, heatmap_size) for b in range(B)])
        w = np.zeros((B, K, 1), dtype=np.float32)
        w[:, :n_valid] = 1.0
        t[:, n_valid:] = 0.0
        target_s.append(t)
        weight_s.append(w)
        kpts_s.append(xy)
    g_of = [b - b % max(1, int(shared_support)) for b in range(B)]     # first row of each row's support group
    if shared_support > 1:
        for j in range(shots):
            for arr_ in (img_s[j], target_s[j], weight_s[j], kpts_s[j]):
                arr_[:] = arr_[g_of]
    img_metas = []
    group_edges = {}
    for b in range(B):
        if g_of[b] not in group_edges:
            group_edges[g_of[b]] = random_skeleton(rng, K, n_valid, skeleton)
        edges = group_edges[g_of[b]]
        img_metas.append(dict(
            sample_skeleton=[edges], query_skeleton=edges,
            query_center=np.array([R / 2.0, R / 2.0], dtype=np.float32),
            query_scale=np.array([R / 200.0, R / 200.0], dtype=np.float32),
            query_image_file=f"synthetic_q_{seed}_{b}.png",
            sample_image_file=[f"synthetic_s_{seed}_{g_of[b]}_{j}.png" for j in range(shots)],
            query_bbox_score=1.0, bbox_id=b,
            # the support crop (the pipeline's Collect meta_keys center / scale / rotation, configs/test/1shot_split1.py:129)
            sample_center=[np.array([R / 2.0, R / 2.0], dtype=np.float32) for _ in range(shots)],
            sample_scale=[np.array([R / 200.0, R / 200.0], dtype=np.float32) for _ in range(shots)],
            sample_rotation=[0 for _ in range(shots)],
            sample_joints_3d=[kpts_s[j][b] for j in range(shots)],
        ))

    def tt(a):
        t = torch.from_numpy(a)
        return t.pin_memory() if pin_memory else t

    return dict(
        img_s=[tt(a) for a in img_s], img_q=tt(img_q),
        target_s=[tt(a) for a in target_s], target_weight_s=[tt(a) for a in weight_s],
        target_q=None, target_weight_q=None, img_metas=img_metas)
