"""Test-loop entry points mirroring /root/reference/EdgeCape/apis/test.py:33-75 (`single_gpu_test`).

The reference calls `model(return_loss=False, **data)` once per batch and waits for it.  Here consecutive batches are
software-pipelined through `EdgeCape.forward_test_async`: while batch i runs its head and its results travel to the
host, the images of batch i+1 are already being copied and its backbone runs beside them.  Results are returned in
order, one dict per batch, exactly what the reference's loop collects.
"""
import collections

import torch

_DATA_KEYS = ("img_s", "target_s", "target_weight_s", "img_q", "target_q", "target_weight_q", "img_metas")


def iter_results(model, batches, depth=None):
    """Yield forward_test's result dict for every batch of `batches` (dicts with the reference's data keys)."""
    model = getattr(model, "module", model)
    model.eval()
    depth = int(depth or model.test_cfg.get("pipeline_depth", 2))
    pending = collections.deque()
    with torch.no_grad():
        for data in batches:
            kw = {k: v for k, v in data.items() if k != "return_loss"}
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
