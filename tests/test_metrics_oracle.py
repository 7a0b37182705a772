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


def test_known_answer_vectors_of_mmpose_unit_tests():
    """Known-answer vectors of mmpose 0.29's own unit tests for these functions
    (tests/test_evaluation/test_top_down_eval.py: test_keypoint_pck_accuracy, test_keypoint_auc), restated here because
    mmpose is absent offline; every expected value is also derivable by hand from the documented definitions
    (PCK = share of valid keypoints whose normalised distance is < thr, per keypoint channel, -1 for a channel with no
    valid sample; AUC = mean PCK over thr = i / num_step).  They pin oracle/metrics_oracle.py -- and through it
    ec_metrics_accumulate (tests/test_ops_gpu.py) -- to mmpose's published behaviour, not merely to itself."""
    # --- test_keypoint_pck_accuracy: 2 samples x 5 keypoints, normaliser 10, thr 0.5
    output, target = np.zeros((2, 5, 2)), np.zeros((2, 5, 2))
    mask = np.array([[True, True, False, True, True], [True, True, False, True, True]])
    thr = np.full((2, 2), 10, dtype=np.float32)
    output[0, 0], target[0, 0] = [10, 0], [10, 0]
    output[0, 1], target[0, 1] = [20, 20], [10, 10]      # distance sqrt(2) > 0.5: miss; sample 1 (zeros) hits -> 0.5
    output[0, 2], target[0, 2] = [0, 0], [-1, 0]         # masked channel -> -1
    output[0, 3], target[0, 3] = [30, 30], [30, 30]
    output[0, 4], target[0, 4] = [0, 10], [0, 10]
    acc, avg_acc, cnt = mo.keypoint_pck_accuracy(output, target, mask, 0.5, thr)
    np.testing.assert_array_almost_equal(acc, np.array([1, 0.5, -1, 1, 1]), decimal=4)
    assert abs(avg_acc - 0.875) < 1e-4 and cnt == 4
    # --- test_keypoint_auc: 1 sample, normaliser 20, 4 steps -> thresholds 0, .25, .5, .75
    output, target = np.zeros((1, 5, 2)), np.zeros((1, 5, 2))
    mask = np.array([[True, True, False, True, True]])
    output[0, 0], target[0, 0] = [10, 4], [10, 5]        # 1 / 20  = 0.05
    output[0, 1], target[0, 1] = [10, 18], [10, 10]      # 8 / 20  = 0.40
    output[0, 2], target[0, 2] = [0, 0], [0, -1]
    output[0, 3], target[0, 3] = [40, 40], [30, 30]      # 14.14 / 20 = 0.707
    output[0, 4], target[0, 4] = [20, 10], [0, 10]       # 20 / 20 = 1.0
    assert abs(mo.keypoint_auc(output, target, mask, 20, 4) - 0.375) < 1e-4      # (0 + 1/4 + 2/4 + 3/4) / 4
    # EPE and NME of the same sample, by hand: mean of (1, 8, 14.1421, 20) and of those / 20
    assert abs(mo.keypoint_epe(output, target, mask) - (1 + 8 + 200 ** 0.5 + 20) / 4) < 1e-4
    assert abs(mo.keypoint_nme(output, target, mask, np.full((1, 2), 20.0)) - (1 + 8 + 200 ** 0.5 + 20) / 80) < 1e-5
