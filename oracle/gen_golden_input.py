"""TEST INFRASTRUCTURE: golden vectors for the input stage, produced by the UNMODIFIED reference.

Runs /root/reference/EdgeCape/datasets/pipelines/top_down_transform.py (`TopDownAffineFewShot`,
`TopDownGenerateTargetFewShot`) and the vendored `post_transforms.get_affine_transform` on seeded synthetic samples,
with OpenCV's `cv2.warpAffine` as installed here (opencv-python 4.13), followed by mmpose 0.29's `ToTensor` /
`NormalizeTensor` (torchvision `to_tensor` + `normalize`; mmpose is absent, these two one-liners are restated).
Output: tests/golden/input_stage.npz.  Run from the repo root:  python -m oracle.gen_golden_input
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "input_stage.npz")


def load_reference_pipeline():
    pipe_dir = os.path.join(REFERENCE_ROOT, "EdgeCape", "datasets", "pipelines")

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    for pkg in ("_ecref", "_ecref.pipelines"):
        m = types.ModuleType(pkg)
        m.__path__ = [pipe_dir]
        sys.modules[pkg] = m
    post = load("_ecref.pipelines.post_transforms", os.path.join(pipe_dir, "post_transforms.py"))

    class _Reg:
        def register_module(self, *a, **k):
            return lambda cls: cls

    def _nyi(*a, **k):
        raise NotImplementedError

    saved = {k: sys.modules.get(k) for k in ("mmcv", "mmpose", "mmpose.datasets", "mmpose.datasets.builder",
                                             "mmpose.core", "mmpose.core.post_processing")}
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mod("mmcv", fileio=types.SimpleNamespace())
    mod("mmpose"); mod("mmpose.datasets"); mod("mmpose.datasets.builder", PIPELINES=_Reg())
    mod("mmpose.core")
    mod("mmpose.core.post_processing", affine_transform=post.affine_transform, fliplr_joints=_nyi,
        get_affine_transform=post.get_affine_transform, get_warp_matrix=_nyi, warp_affine_joints=_nyi)
    tdt = load("_ecref.pipelines.top_down_transform", os.path.join(pipe_dir, "top_down_transform.py"))
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    return post, tdt


def make_sample(rng, Hs, Ws, K, R):
    img = rng.integers(0, 256, (Hs, Ws, 3), dtype=np.uint8)
    # bbox -> centre / scale as test_dataset.py:224-252 (_xywh2cs, padding 1.25, pixel_std 200)
    x, y = rng.uniform(0, Ws * 0.3), rng.uniform(0, Hs * 0.3)
    w, h = rng.uniform(Ws * 0.4, Ws * 0.7), rng.uniform(Hs * 0.4, Hs * 0.7)
    center = np.array([x + w * 0.5, y + h * 0.5], dtype=np.float32)
    aspect = 1.0
    if w > aspect * h:
        h = w / aspect
    elif w < aspect * h:
        w = h * aspect
    scale = np.array([w / 200.0, h / 200.0], dtype=np.float32) * 1.25
    joints = np.zeros((K, 3), dtype=np.float32)
    joints[:, 0] = rng.uniform(x - 10, x + w + 10, K)
    joints[:, 1] = rng.uniform(y - 10, y + h + 10, K)
    vis = np.zeros((K, 3), dtype=np.float32)
    v = (rng.uniform(size=K) > 0.2).astype(np.float32)
    vis[:, 0] = v
    vis[:, 1] = v
    return img, center, scale, joints, vis


def main():
    post, tdt = load_reference_pipeline()
    rng = np.random.default_rng(20260117)
    R, HM, K = 256, 64, 17
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    affine = tdt.TopDownAffineFewShot()
    gen = tdt.TopDownGenerateTargetFewShot(sigma=1)
    out = {}
    sizes = [(97, 131), (240, 320), (300, 200), (64, 64)]
    for i, (Hs, Ws) in enumerate(sizes):
        img, c, s, joints, vis = make_sample(rng, Hs, Ws, K, R)
        ann = dict(image_size=np.array([R, R]), heatmap_size=np.array([HM, HM]), joint_weights=None,
                   use_different_joint_weights=False)
        res = dict(img=img.copy(), joints_3d=joints.copy(), joints_3d_visible=vis.copy(), center=c, scale=s, rotation=0,
                   ann_info=ann)
        trans = post.get_affine_transform(c, s, 0, ann["image_size"])
        res = affine(res)
        warped = res["img"]
        t = torch.from_numpy(warped.transpose(2, 0, 1)).contiguous().to(torch.float32).div(255)      # ToTensor
        t = t.sub(torch.tensor(mean).view(3, 1, 1)).div(torch.tensor(std).view(3, 1, 1))              # NormalizeTensor
        target, weight = gen._msra_generate_target(ann, res["joints_3d"], res["joints_3d_visible"], 1)
        out.update({f"img{i}": img, f"center{i}": c, f"scale{i}": s, f"joints{i}": joints, f"vis{i}": vis,
                    f"trans{i}": trans, f"warped{i}": warped, f"tensor_sum{i}": np.array(t.double().sum().item()),
                    f"tensor_corner{i}": t[:, :4, :4].numpy(),
                    f"joints_t{i}": res["joints_3d"], f"target{i}": target, f"weight{i}": weight})
    out["n"] = np.array(len(sizes))
    out["mean"] = np.array(mean, dtype=np.float32)
    out["std"] = np.array(std, dtype=np.float32)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
