"""The device input stage (edgecape_b200/inputs.py) against the golden vectors of the UNMODIFIED reference pipeline +
OpenCV (tests/golden/input_stage.npz).  CPU: host geometry and the orchestration with the C ABI emulated; GPU: the
kernels themselves -- the crop must reproduce cv2.warpAffine bit for bit."""
import os

import numpy as np
import pytest
import torch

from edgecape_b200 import inputs

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "input_stage.npz"))
R, HM = 256, 64


def _want_tensor(i):
    t = torch.from_numpy(G[f"warped{i}"].transpose(2, 0, 1)).contiguous().to(torch.float32).div(255)
    return t.sub(torch.tensor(G["mean"]).view(3, 1, 1)).div(torch.tensor(G["std"]).view(3, 1, 1))


def test_host_geometry_matches_reference():
    for i in range(int(G["n"])):
        M = inputs.get_affine_transform(G[f"center{i}"], G[f"scale{i}"], 0, (R, R))
        assert np.allclose(M, G[f"trans{i}"], rtol=0, atol=1e-9)
        jt = inputs.affine_transform_joints(G[f"joints{i}"], G[f"vis{i}"], G[f"trans{i}"])
        assert np.allclose(jt, G[f"joints_t{i}"], atol=1e-4)
    c, s = inputs.xywh2cs(10.0, 20.0, 100.0, 50.0, (256, 256))
    assert np.allclose(c, [60.0, 45.0]) and np.allclose(s, [100 / 200 * 1.25, 100 / 200 * 1.25])


def _run(dev):
    stage = inputs.InputStage((R, R), (HM, HM), 1, G["mean"], G["std"])
    for i in range(int(G["n"])):
        img = torch.from_numpy(G[f"img{i}"]).to(dev)
        got = stage.crop_normalize(img, G[f"trans{i}"])
        assert torch.equal(got.cpu(), _want_tensor(i)), f"crop {i}: max diff {(got.cpu() - _want_tensor(i)).abs().max()}"
        j = torch.from_numpy(np.ascontiguousarray(G[f"joints_t{i}"])).to(dev)          # [K,3]: ld 3
        v = torch.from_numpy(np.ascontiguousarray(G[f"vis{i}"])).to(dev)
        target, weight = stage.targets(j, v)
        assert torch.equal(weight.cpu(), torch.from_numpy(G[f"weight{i}"]))
        assert np.allclose(target.cpu().numpy(), G[f"target{i}"], rtol=0, atol=1.2e-7)
    # batched targets [B,K,2] and the whole-sample helper
    jb = torch.from_numpy(np.stack([G[f"joints_t{i}"][:, :2] for i in range(2)])).contiguous().to(dev)
    vb = torch.from_numpy(np.stack([G[f"vis{i}"][:, :1] for i in range(2)])).contiguous().to(dev)
    tb, wb = stage.targets(jb, vb)
    assert tuple(tb.shape) == (2, jb.shape[1], HM, HM) and tuple(wb.shape) == (2, jb.shape[1], 1)
    assert np.allclose(tb[1].cpu().numpy(), G["target1"], rtol=0, atol=1.2e-7)


def test_input_stage_host_orchestration(monkeypatch):
    from tests import cpu_emulator
    cpu_emulator.install(monkeypatch)
    _run(torch.device("cpu"))


@pytest.mark.gpu
def test_input_stage_kernels_bit_exact_crop():
    _run(torch.device("cuda"))


@pytest.mark.gpu
def test_crop_kernel_matches_oracle_for_rotated_transforms():
    """The oracle's fixed-point warp is pinned against OpenCV (tests/test_input_oracle.py); here the kernel is compared
    with the oracle on rotated / anisotropic transforms and odd image sizes."""
    from oracle import input_oracle as io
    rng = np.random.default_rng(9)
    dev = torch.device("cuda")
    for _ in range(5):
        Hs, Ws = int(rng.integers(30, 300)), int(rng.integers(30, 300))
        img = rng.integers(0, 256, (Hs, Ws, 3), dtype=np.uint8)
        ang = np.deg2rad(rng.uniform(-90, 90))
        sx, sy = rng.uniform(0.2, 3.0, 2)
        M = np.array([[sx * np.cos(ang), -sy * np.sin(ang), rng.uniform(-50, 50)],
                      [sx * np.sin(ang), sy * np.cos(ang), rng.uniform(-50, 50)]])
        W, H = int(rng.integers(8, 130)), int(rng.integers(8, 130))
        stage = inputs.InputStage((W, H), (HM, HM), 1, G["mean"], G["std"])
        got = stage.crop_normalize(torch.from_numpy(img).to(dev), M).cpu().numpy()
        want = io.to_tensor_normalize(io.warp_affine_u8(img, M, W, H), G["mean"], G["std"])
        assert np.array_equal(got, want), float(np.abs(got - want).max())


@pytest.mark.gpu
def test_gather_blocks_and_empty_batches():
    from edgecape_b200 import ops
    from edgecape_b200.parallel import new_metric_counters
    dev = torch.device("cuda")
    src = torch.randn(5, 7, 12, device=dev)
    idx = torch.tensor([4, 0, 0, 3, 1, 4], dtype=torch.int32, device=dev)
    assert torch.equal(ops.gather_blocks(src, idx), src[idx.long()])
    tok = torch.randn(4, 9, 10, device=dev)                       # strided view (cls row dropped), odd sizes -> scalar path
    idx2 = torch.tensor([3, 0, 2], dtype=torch.int32, device=dev)
    assert torch.equal(ops.gather_blocks(tok[:, 1:, :], idx2), tok[:, 1:, :][idx2.long()])
    c = new_metric_counters(dev)
    z = torch.zeros(0, 5, 2, device=dev)
    ops.metrics_accumulate_(c, z, z, torch.zeros(0, 5, dtype=torch.uint8, device=dev), torch.zeros(0, 2, device=dev),
                            torch.tensor([0.2], device=dev))
    assert float(c.sum()) == 0.0
