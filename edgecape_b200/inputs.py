"""Input stage on the device: the reference's per-sample test pipeline
(/root/reference/configs/test/1shot_split1.py:99-117) --

    TopDownAffineFewShot  ->  ToTensor  ->  NormalizeTensor  ->  TopDownGenerateTargetFewShot(sigma)

(/root/reference/EdgeCape/datasets/pipelines/top_down_transform.py:18-67, 70-199; mmpose 0.29 ToTensor / NormalizeTensor)
-- with the image crop, the normalisation and the heat-map targets computed by CUDA kernels, so raw uint8 images and
keypoint lists (a few KB) cross PCIe instead of fp32 crops and dense 64x64 heat-maps (51 MB per 16-query step).

The tiny per-sample geometry (bbox -> centre/scale, the 2x3 affine matrix, mapping the keypoints through it) stays on the
host in float64 exactly as the reference computes it.
"""
import ctypes

import numpy as np
import torch

from . import _lib, ops


def xywh2cs(x, y, w, h, image_size, padding=1.25, pixel_std=200.0):
    """datasets/datasets/mp100/test_dataset.py:224-252 (`_xywh2cs`): bbox -> centre, scale (aspect-corrected, padded)."""
    aspect = image_size[0] / image_size[1]
    center = np.array([x + w * 0.5, y + h * 0.5], dtype=np.float32)
    if w > aspect * h:
        h = w * 1.0 / aspect
    elif w < aspect * h:
        w = h * aspect
    scale = np.array([w / pixel_std, h / pixel_std], dtype=np.float32)
    return center, scale * padding


def get_affine_transform(center, scale, rot, output_size):
    """post_transforms.py:10-64 (shift 0): the double 2x3 matrix source -> crop, as cv2.getAffineTransform solves it."""
    scale_tmp = np.asarray(scale, dtype=np.float32) * 200.0
    src_w = scale_tmp[0]
    dst_w, dst_h = float(output_size[0]), float(output_size[1])
    r = np.pi * rot / 180
    sn, cs = np.sin(r), np.cos(r)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0] = center
    src[1] = np.asarray(center) + np.array([-(src_w * -0.5) * sn, (src_w * -0.5) * cs])
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = dst[0] + np.array([0.0, dst_w * -0.5], dtype=np.float32)
    for p in (src, dst):
        d = p[0] - p[1]
        p[2] = p[1] + np.array([-d[1], d[0]], dtype=np.float32)
    A = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        A[i, 0:2], A[i, 2] = src[i], 1.0
        A[i + 3, 3:5], A[i + 3, 5] = src[i], 1.0
        b[i], b[i + 3] = dst[i, 0], dst[i, 1]
    return np.linalg.solve(A, b).reshape(2, 3)


def affine_transform_joints(joints, visible, trans):
    """top_down_transform.py:58-61: visible joints go through the matrix (float64), invisible ones stay."""
    out = np.array(joints, dtype=np.float32, copy=True)
    vis = np.asarray(visible)[:, 0] > 0.0
    pts = np.concatenate((np.asarray(joints, dtype=np.float64)[:, :2], np.ones((len(out), 1))), axis=1)
    out[vis, 0:2] = (pts @ np.asarray(trans, dtype=np.float64).T)[vis]
    return out


class InputStage:
    """Device-side replacement of the reference's test pipeline for one sample / one batch."""

    def __init__(self, image_size=(256, 256), heatmap_size=(64, 64), sigma=1, mean=(0.485, 0.456, 0.406),
                 std=(0.229, 0.224, 0.225)):
        self.image_size = (int(image_size[0]), int(image_size[1]))          # (W, H)
        self.heatmap_size = (int(heatmap_size[0]), int(heatmap_size[1]))    # (W, H)
        self.sigma = float(sigma)
        self._mean = (ctypes.c_float * 3)(*mean)
        self._std = (ctypes.c_float * 3)(*std)

    def crop_normalize(self, img_u8, trans, out=None):
        """img_u8: uint8 [Hs, Ws, 3] CUDA tensor (RGB as LoadImageFromFile yields it); trans: 2x3 float64 (host).
        Returns the fp32 [3, H, W] network input."""
        ops._chk(img_u8, "img_u8", torch.uint8)
        assert img_u8.dim() == 3 and img_u8.shape[2] == 3 and img_u8.stride(2) == 1 and img_u8.stride(1) == 3
        W, H = self.image_size
        if out is None:
            out = ops.empty(3, H, W, device=img_u8.device)
        assert out.is_contiguous() and tuple(out.shape) == (3, H, W)
        M = (ctypes.c_double * 6)(*np.asarray(trans, dtype=np.float64).reshape(-1))
        _lib.call("ec_warp_affine_normalize_u8", img_u8.data_ptr(), img_u8.shape[0], img_u8.shape[1], img_u8.stride(0),
                  ctypes.addressof(M), out.data_ptr(), H, W, ctypes.addressof(self._mean), ctypes.addressof(self._std),
                  ops._stream())
        return out

    def targets(self, joints, visible):
        """joints [..., K, >=2] crop-pixel coordinates, visible [..., K, >=1] (CUDA fp32) ->
        target [..., K, Hh, Wh], target_weight [..., K, 1]."""
        ops._chk(joints, "joints"); ops._chk(visible, "visible")
        assert joints.is_contiguous() and visible.is_contiguous()
        lead = tuple(joints.shape[:-1])
        n = int(np.prod(lead))
        Wh, Hh = self.heatmap_size
        target = ops.empty(*lead, Hh, Wh, device=joints.device)
        weight = ops.empty(*lead, 1, device=joints.device)
        _lib.call("ec_msra_targets", joints.data_ptr(), joints.shape[-1], visible.data_ptr(), visible.shape[-1],
                  target.data_ptr(), weight.data_ptr(), n, self.image_size[0], self.image_size[1], Wh, Hh, self.sigma,
                  ops._stream())
        return target, weight

    def sample(self, img_u8, bbox, joints_3d, joints_3d_visible, rotation=0.0):
        """One sample end to end: bbox [x, y, w, h], joints_3d [K,3] / joints_3d_visible [K,3] in source pixels (numpy).
        Returns dict(img [3,H,W], target [K,Hh,Wh], target_weight [K,1], center, scale, joints_3d (crop pixels))."""
        center, scale = xywh2cs(*bbox, self.image_size)
        trans = get_affine_transform(center, scale, rotation, self.image_size)
        jt = affine_transform_joints(joints_3d, joints_3d_visible, trans)
        dev = img_u8.device
        img = self.crop_normalize(img_u8, trans)
        j = torch.from_numpy(np.ascontiguousarray(jt[:, :2])).to(dev, non_blocking=True)
        v = torch.from_numpy(np.ascontiguousarray(np.asarray(joints_3d_visible, dtype=np.float32)[:, :1])).to(dev, non_blocking=True)
        target, weight = self.targets(j, v)
        return dict(img=img, target=target, target_weight=weight, center=center, scale=scale, joints_3d=jt)
