"""TEST INFRASTRUCTURE (oracle): the evaluation metrics of the reference's test loop, restated on the CPU.

Path: /root/reference/EdgeCape/datasets/datasets/mp100/test_base_dataset.py:71-155 (`_report_metric`) calls
`keypoint_pck_accuracy`, `keypoint_nme`, `keypoint_auc`, `keypoint_epe` of **mmpose 0.29.0**
(`mmpose/core/evaluation/top_down_eval.py`; pinned by the reference's README.md:66, not vendored under
/root/reference), one sample at a time (N = 1), and averages the per-sample values.  The four functions are restated
here from their published algorithm; `report_metric` is the reference's reduction.  Only tests/, smoke() and bench.py's
CPU legs may import this module.

Pinning: mmpose itself is absent offline, so this module cannot be executed against it.  It is pinned to the
known-answer vectors of mmpose 0.29's own unit tests for these functions (test_keypoint_pck_accuracy: acc = [1, 0.5,
-1, 1, 1], mean 0.875, 4 valid channels; test_keypoint_auc: 0.375), restated in
tests/test_metrics_oracle.py::test_known_answer_vectors_of_mmpose_unit_tests together with hand-derived EPE / NME
values for the same sample -- "partially pinned": known answers, not outputs of mmpose run here.
`_keypoint_pck_accuracy` in oracle/ref_shims.py is the same restatement the shimmed reference runs with.
"""
import numpy as np


def _calc_distances(preds, targets, mask, normalize):
    """mmpose top_down_eval._calc_distances: [K, N] normalised distances, -1 where masked; samples whose normaliser
    has a zero are masked out entirely, non-positive normalisers become 1e6."""
    N, K, _ = preds.shape
    _mask = mask.copy()
    _mask[np.where((normalize == 0).sum(1))[0], :] = False
    distances = np.full((N, K), -1, dtype=np.float32)
    normalize = normalize.copy()
    normalize[np.where(normalize <= 0)] = 1e6
    distances[_mask] = np.linalg.norm(((preds - targets) / normalize[:, None, :])[_mask], axis=-1)
    return distances.T


def _distance_acc(distances, thr):
    valid = distances != -1
    n = valid.sum()
    return (distances[valid] < thr).sum() / n if n > 0 else -1


def keypoint_pck_accuracy(pred, gt, mask, thr, normalize):
    distances = _calc_distances(pred, gt, mask, normalize)
    acc = np.array([_distance_acc(d, thr) for d in distances])
    valid_acc = acc[acc >= 0]
    cnt = len(valid_acc)
    return acc, (valid_acc.mean() if cnt > 0 else 0), cnt


def keypoint_nme(pred, gt, mask, normalize_factor):
    distances = _calc_distances(pred, gt, mask, normalize_factor)
    dv = distances[distances != -1]
    return dv.sum() / max(1, len(dv))


def keypoint_auc(pred, gt, mask, normalize, num_step=20):
    nor = np.tile(np.array([[normalize, normalize]]), (pred.shape[0], 1))
    x = [1.0 * i / num_step for i in range(num_step)]
    y = [keypoint_pck_accuracy(pred, gt, mask, thr, nor)[1] for thr in x]
    auc = 0
    for i in range(num_step):
        auc += 1.0 / num_step * y[i]
    return auc


def keypoint_epe(pred, gt, mask):
    distances = _calc_distances(pred, gt, mask, np.ones((pred.shape[0], pred.shape[2]), dtype=np.float32))
    dv = distances[distances != -1]
    return dv.sum() / max(1, len(dv))


def report_metric(outputs, gts, masks, bbox_thr, thresholds=(0.05, 0.1, 0.15, 0.2, 0.25)):
    """test_base_dataset.py:119-154: per-sample metrics, then their means.  outputs / gts [B,K,2], masks bool [B,K],
    bbox_thr [B] (= max(bbox_w, bbox_h), :111-113).  Returns dict incl. the per-sample sums the device kernel
    accumulates."""
    B = len(outputs)
    pck = {t: [] for t in thresholds}
    nme, auc, epe = [], [], []
    for o, g, m, thr in zip(outputs, gts, masks, bbox_thr):
        o1, g1, m1 = np.expand_dims(o, 0), np.expand_dims(g, 0), np.expand_dims(m, 0)
        tb = np.expand_dims(np.array([thr, thr]), 0)
        for t in thresholds:
            pck[t].append(keypoint_pck_accuracy(o1, g1, m1, t, tb)[1])
        nme.append(keypoint_nme(o1, g1, m1, tb))
        auc.append(keypoint_auc(o1, g1, m1, tb[0, 0]))
        epe.append(keypoint_epe(o1, g1, m1))
    out = {f"PCK@{t}": float(np.mean(pck[t])) for t in thresholds}
    out["mPCK"] = float(np.mean([out[f"PCK@{t}"] for t in thresholds]))
    out.update(NME=float(np.mean(nme)), AUC=float(np.mean(auc)), EPE=float(np.mean(epe)), samples=B)
    out["sums"] = [float(np.sum(pck[t])) for t in thresholds] + [float(np.sum(nme)), float(np.sum(auc)),
                                                                   float(np.sum(epe)), float(B)]
    return out
