"""Test-loop entry points mirroring /root/reference/EdgeCape/apis/test.py:33-75 (`single_gpu_test`).

The reference calls `model(return_loss=False, **data)` once per batch and waits for it.  Here consecutive batches are
software-pipelined through `EdgeCape.forward_test_async`: while batch i runs its head and its results travel to the
host, the images of batch i+1 are already being copied and its backbone runs beside them.  Results are returned in
order, one dict per batch, exactly what the reference's loop collects.
"""
import collections

import torch

from .detector import unwrap_data_container

_DATA_KEYS = ("img_s", "target_s", "target_weight_s", "img_q", "target_q", "target_weight_q", "img_metas")


def iter_results(model, batches, depth=None):
    """Yield forward_test's result dict for every batch of `batches` (dicts with the reference's data keys)."""
    model = getattr(model, "module", model)
    model.eval()
    depth = int(depth or model.test_cfg.get("pipeline_depth", 2))
    pending = collections.deque()
    with torch.no_grad():
        for data in batches:
            kw = {k: unwrap_data_container(v) for k, v in data.items() if k != "return_loss"}
            if getattr(model, "use_cuda_graph", False):
                pending.append(model.forward_test_async(**kw))
            else:
                pending.append(_Done(model.forward_test(**kw)))
            if len(pending) >= depth:
                yield pending.popleft().result()
        while pending:
            yield pending.popleft().result()


class _Done:
    def __init__(self, res):
        self._res = res

    def result(self):
        return self._res


def single_gpu_test(model, data_loader, pck=False, depth=None):
    """apis/test.py:33-50: run the model over the loader, return the list of per-batch result dicts."""
    return list(iter_results(model, data_loader, depth=depth))


def multi_gpu_test(model, data_loader, tmpdir=None, gpu_collect=True, pck=False, depth=None, size=None):
    """apis/test.py:53-90 (`multi_gpu_test`): every rank runs the pipelined loop over its own loader, then the
    per-batch result dicts are gathered on rank 0 and re-interleaved exactly like `collect_results_gpu` /
    `collect_results_cpu` (:93-198: `zip(*parts)`, truncated to the dataset size because the distributed sampler pads).
    Returns the ordered list on rank 0 and None elsewhere.  `tmpdir` / `gpu_collect` are accepted for signature
    compatibility; the exchange is one `all_gather_object`.  (For metrics only, prefer the device-side counters +
    one all-reduce: parallel.new_metric_counters / allreduce_counters.)"""
    import torch.distributed as dist
    part = list(iter_results(model, data_loader, depth=depth))
    if size is None:
        ds = getattr(data_loader, "dataset", None)
        size = len(ds) if ds is not None else None
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return part if size is None else part[:size]
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, part)
    if dist.get_rank() != 0:
        return None
    ordered = []
    for res in zip(*parts):
        ordered.extend(list(res))
    return ordered if size is None else ordered[:size]
