"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's input stage.

Follows /root/reference/EdgeCape/datasets/pipelines/top_down_transform.py:35-67 (`TopDownAffineFewShot`: affine crop
with `cv2.warpAffine(..., INTER_LINEAR)`), :113-199 (`_msra_generate_target`), post_transforms.py:10-112
(`get_affine_transform`, `affine_transform`) and mmpose 0.29 `ToTensor` / `NormalizeTensor` (torchvision `to_tensor`,
`normalize`).  `cv2.warpAffine` is OpenCV (a third-party dependency, not under /root/reference): its 8-bit bilinear
path is restated from the published algorithm -- inverse matrix in double, source coordinates in 10-bit fixed point
rounded to 1/32 pixel, 15-bit bilinear weights, `(sum + 2^14) >> 15` -- and pinned bit-exactly against opencv-python
4.13 through tests/golden/input_stage.npz (oracle/gen_golden_input.py).
Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np


def get_3rd_point(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def get_affine_transform(center, scale, rot, output_size):
    """post_transforms.py:10-64 with shift = 0: the 2x3 matrix mapping source pixels to the crop (double)."""
    scale_tmp = np.asarray(scale, dtype=np.float32) * 200.0
    src_w = scale_tmp[0]
    dst_w, dst_h = float(output_size[0]), float(output_size[1])
    r = np.pi * rot / 180
    sn, cs = np.sin(r), np.cos(r)
    src_dir = np.array([0.0 * cs - (src_w * -0.5) * sn, 0.0 * sn + (src_w * -0.5) * cs])
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0] = center
    src[1] = np.asarray(center) + src_dir
    src[2] = get_3rd_point(src[0], src[1])
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0.0, dst_w * -0.5])
    dst[2] = get_3rd_point(dst[0], dst[1])
    # cv2.getAffineTransform: the 6x6 linear system  [x y 1 0 0 0; 0 0 0 x y 1] m = [u; v]
    A = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        A[i, 0:2], A[i, 2] = src[i], 1.0
        A[i + 3, 3:5], A[i + 3, 5] = src[i], 1.0
        b[i], b[i + 3] = dst[i, 0], dst[i, 1]
    return np.linalg.solve(A, b).reshape(2, 3)


def affine_transform_joints(joints, visible, trans):
    """top_down_transform.py:58-61: visible joints are mapped through the matrix, the others stay."""
    out = joints.copy()
    for i in range(len(joints)):
        if visible[i, 0] > 0.0:
            out[i, 0:2] = (np.array(trans) @ np.array([joints[i, 0], joints[i, 1], 1.0]))[:2]
    return out


def invert_affine(M):
    M = np.asarray(M, dtype=np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    iM = np.zeros((2, 3))
    iM[0, 0], iM[1, 1] = M[1, 1] * D, M[0, 0] * D
    iM[0, 1], iM[1, 0] = M[0, 1] * (-D), M[1, 0] * (-D)
    iM[0, 2] = -iM[0, 0] * M[0, 2] - iM[0, 1] * M[1, 2]
    iM[1, 2] = -iM[1, 0] * M[0, 2] - iM[1, 1] * M[1, 2]
    return iM


def warp_affine_u8(img, M, W, H):
    """cv2.warpAffine(img, M, (W, H), flags=INTER_LINEAR) for uint8 HxWxC, constant (0) border."""
    iM = invert_affine(M)
    AB = 1024.0
    xs = np.arange(W)
    ys = np.arange(H)
    adelta = np.rint(iM[0, 0] * xs * AB).astype(np.int64)
    bdelta = np.rint(iM[1, 0] * xs * AB).astype(np.int64)
    X0 = np.rint((iM[0, 1] * ys + iM[0, 2]) * AB).astype(np.int64) + 16
    Y0 = np.rint((iM[1, 1] * ys + iM[1, 2]) * AB).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx, sy, fx, fy = X >> 5, Y >> 5, X & 31, Y & 31
    Hs, Ws = img.shape[:2]
    src = img.astype(np.int64)

    def px(yy, xx):
        ok = (yy >= 0) & (yy < Hs) & (xx >= 0) & (xx < Ws)
        return src[np.clip(yy, 0, Hs - 1), np.clip(xx, 0, Ws - 1)] * ok[..., None]

    w00, w01 = 32 * (32 - fx) * (32 - fy), 32 * fx * (32 - fy)
    w10, w11 = 32 * (32 - fx) * fy, 32 * fx * fy
    acc = (px(sy, sx) * w00[..., None] + px(sy, sx + 1) * w01[..., None] + px(sy + 1, sx) * w10[..., None] +
           px(sy + 1, sx + 1) * w11[..., None])
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def to_tensor_normalize(img_u8, mean, std):
    """mmpose ToTensor + NormalizeTensor: HWC uint8 -> CHW float32, /255, (x - mean) / std (all fp32)."""
    t = img_u8.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    return (t - np.asarray(mean, np.float32).reshape(3, 1, 1)) / np.asarray(std, np.float32).reshape(3, 1, 1)


def msra_targets(joints, visible, image_size, heatmap_size, sigma):
    """top_down_transform.py:113-199, `unbiased_encoding=False`: joints [K,>=2] in crop pixels, visible [K,>=1] ->
    target [K,H,W] (un-normalised Gaussian patches, clipped at the border), target_weight [K,1]."""
    K = len(joints)
    W, H = int(heatmap_size[0]), int(heatmap_size[1])
    target = np.zeros((K, H, W), dtype=np.float32)
    weight = np.zeros((K, 1), dtype=np.float32)
    tmp = sigma * 3
    stride = np.asarray(image_size, dtype=np.float64) / np.array([W, H], dtype=np.float64)
    size = 2 * tmp + 1
    ax = np.arange(0, size, 1, np.float32)
    g = np.exp(-((ax[None, :] - size // 2) ** 2 + (ax[:, None] - size // 2) ** 2) / (2 * sigma ** 2))
    for k in range(K):
        weight[k] = visible[k, 0]
        mx = int(joints[k][0] / stride[0] + 0.5)
        my = int(joints[k][1] / stride[1] + 0.5)
        ul = [int(mx - tmp), int(my - tmp)]
        br = [int(mx + tmp + 1), int(my + tmp + 1)]
        if ul[0] >= W or ul[1] >= H or br[0] < 0 or br[1] < 0:
            weight[k] = 0
        if weight[k] > 0.5:
            gx = max(0, -ul[0]), min(br[0], W) - ul[0]
            gy = max(0, -ul[1]), min(br[1], H) - ul[1]
            ix = max(0, ul[0]), min(br[0], W)
            iy = max(0, ul[1]), min(br[1], H)
            target[k][iy[0]:iy[1], ix[0]:ix[1]] = g[gy[0]:gy[1], gx[0]:gx[1]]
    return target, weight
