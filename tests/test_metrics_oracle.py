"""CPU: the restated mmpose metrics (oracle/metrics_oracle.py) on hand-checkable cases, the emulated
ec_metrics_accumulate against them, and the reduction helpers."""
import numpy as np
import torch

from oracle import metrics_oracle as mo


def test_metric_restatement_on_a_hand_case():
    # one sample, 4 keypoints, bbox normaliser 100: distances 0, 10, 30, (masked)
    pred = np.array([[[0.0, 0.0], [10.0, 0.0], [0.0, 30.0], [500.0, 500.0]]])
    gt = np.zeros((1, 4, 2))
    mask = np.array([[True, True, True, False]])
    nor = np.array([[100.0, 100.0]])
    assert mo.keypoint_pck_accuracy(pred, gt, mask, 0.2, nor)[1] == 2 / 3
    assert mo.keypoint_pck_accuracy(pred, gt, mask, 0.1, nor)[1] == 1 / 3          # strict <
    assert abs(mo.keypoint_nme(pred, gt, mask, nor) - (0.0 + 0.1 + 0.3) / 3) < 1e-7
    assert abs(mo.keypoint_epe(pred, gt, mask) - 40.0 / 3) < 1e-5
    # AUC: thresholds 0, .05, ..., .95: d=0 counts for 19 of them, d=.1 for 17 (thr > .1), d=.3 for 13 (thr > .3)
    assert abs(mo.keypoint_auc(pred, gt, mask, 100.0) - (19 + 17 + 13) / 3 / 20) < 1e-7
    # no valid keypoint -> every metric 0; zero normaliser masks PCK / NME but not EPE
    none = np.zeros((1, 4), dtype=bool)
    assert mo.keypoint_pck_accuracy(pred, gt, none, 0.2, nor)[1] == 0 and mo.keypoint_nme(pred, gt, none, nor) == 0
    z = np.array([[0.0, 0.0]])
    assert mo.keypoint_pck_accuracy(pred, gt, mask, 0.2, z)[1] == 0 and mo.keypoint_epe(pred, gt, mask) > 0


def test_emulated_metrics_kernel_and_summary(monkeypatch):
    from tests import cpu_emulator
    from edgecape_b200 import ops
    from edgecape_b200.parallel import PCK_THRESHOLDS, new_metric_counters, summarize_metrics
    cpu_emulator.install(monkeypatch)
    B, K = 5, 12
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(B, K, 2, generator=g) * 100
    pred = gt + torch.randn(B, K, 2, generator=g) * 15
    valid = (torch.rand(B, K, generator=g) > 0.25)
    bbox = torch.tensor([100.0, 60.0, 80.0, 120.0, 90.0])
    c = new_metric_counters("cpu")
    ops.metrics_accumulate_(c, pred, gt, valid.to(torch.uint8), torch.stack((bbox, bbox), 1).contiguous(),
                            torch.tensor(PCK_THRESHOLDS, dtype=torch.float32))
    want = mo.report_metric(list(pred.double().numpy()), list(gt.double().numpy()), list(valid.numpy()),
                            list(bbox.double().numpy()), PCK_THRESHOLDS)
    got = summarize_metrics(c)
    for k in ("PCK@0.1", "mPCK", "NME", "AUC", "EPE"):
        assert abs(got[k] - want[k]) < 1e-6 * max(1.0, abs(want[k])), k
