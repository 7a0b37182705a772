"""Multi-GPU plumbing for the evaluation loop: queries are independent units, so the path shards
with NO data-path collective; the only exchange is one all-reduce(SUM) of the fp64 PCK counters at
the end.  It replaces the reference's pickled `all_gather` of per-sample
results (/root/reference/EdgeCape/apis/test.py:154-198); the reference shards with a
DistributedSampler and re-interleaves (:193-196), here every rank owns a contiguous slice (order is
irrelevant for a mean of per-sample ratios).
"""
import torch
import torch.distributed as dist

PCK_THRESHOLDS = (0.05, 0.1, 0.15, 0.2, 0.25)      # datasets/datasets/mp100/test_base_dataset.py PCK_threshold_list


def shard_range(n_items, rank, world_size):
    """Contiguous slice [lo, hi) of `n_items` owned by `rank` (sizes differ by at most one)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return n_items * rank // world_size, n_items * (rank + 1) // world_size


def new_counters(device, n_thresholds=len(PCK_THRESHOLDS)):
    """[sum of per-sample PCK@t for each t ..., n_samples] in fp64 (ec_pck_accumulate's layout)."""
    return torch.zeros(n_thresholds + 1, dtype=torch.float64, device=device)


def allreduce_counters(counters):
    """The single collective of the path.  No-op without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def summarize_pck(counters, thresholds=PCK_THRESHOLDS):
    """Mean of per-sample PCK per threshold + mPCK, exactly `_report_metric`'s reduction
    (test_base_dataset.py:119-133)."""
    c = counters.detach().cpu().tolist()
    n = max(c[len(thresholds)], 1.0)
    out = {f"PCK@{t}": c[i] / n for i, t in enumerate(thresholds)}
    out["mPCK"] = sum(out.values()) / len(thresholds)
    out["samples"] = int(c[len(thresholds)])
    return out


def new_metric_counters(device, n_thresholds=len(PCK_THRESHOLDS)):
    """[sum PCK@t ..., sum NME, sum AUC, sum EPE, n_samples] in fp64 (ec_metrics_accumulate's layout)."""
    return torch.zeros(n_thresholds + 4, dtype=torch.float64, device=device)


def summarize_metrics(counters, thresholds=PCK_THRESHOLDS):
    """`_report_metric`'s means (test_base_dataset.py:119-154) from the all-reduced counter vector."""
    c = counters.detach().cpu().tolist()
    T = len(thresholds)
    n = max(c[T + 3], 1.0)
    out = {f"PCK@{t}": c[i] / n for i, t in enumerate(thresholds)}
    out["mPCK"] = sum(out.values()) / T
    out.update(NME=c[T] / n, AUC=c[T + 1] / n, EPE=c[T + 2] / n, samples=int(c[T + 3]))
    return out
