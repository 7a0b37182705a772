"""CPU, world_size 2 over gloo: the N>1 path -- contiguous query shards per rank and ONE all-reduce
of the fp64 PCK counters -- gives exactly the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from edgecape_b200.parallel import PCK_THRESHOLDS, allreduce_counters, new_counters, shard_range, summarize_pck


def _per_sample_counters(pred, gt, valid, norm):
    """numpy statement of mmpose keypoint_pck_accuracy with N = 1 per sample (what ec_pck_accumulate does)."""
    c = np.zeros(len(PCK_THRESHOLDS) + 1)
    for b in range(pred.shape[0]):
        d = np.sqrt((((pred[b] - gt[b]) / norm[b]) ** 2).sum(-1))
        for t, thr in enumerate(PCK_THRESHOLDS):
            if valid[b].any():
                c[t] += (d[valid[b]] < thr).mean()
        c[-1] += 1
    return c


def _data(n=37, K=17):
    rng = np.random.default_rng(0)
    gt = rng.uniform(0, 200, (n, K, 2))
    pred = gt + rng.normal(0, 25, (n, K, 2))
    valid = rng.uniform(size=(n, K)) > 0.3
    valid[5] = False
    norm = np.full((n, 2), 200.0)
    return pred, gt, valid, norm


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pred, gt, valid, norm = _data()
    lo, hi = shard_range(pred.shape[0], rank, world)
    c = new_counters("cpu")
    c += torch.from_numpy(_per_sample_counters(pred[lo:hi], gt[lo:hi], valid[lo:hi], norm[lo:hi]))
    allreduce_counters(c)
    if rank == 0:
        q.put(c.tolist())
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 16, 37, 128):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_two_rank_counters_equal_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _per_sample_counters(*_data())
    assert np.allclose(got, want, atol=1e-12)
    s_ = summarize_pck(torch.tensor(got, dtype=torch.float64))
    assert s_["samples"] == 37 and abs(s_["PCK@0.2"] - want[3] / 37) < 1e-12
    assert abs(s_["mPCK"] - want[:5].sum() / 37 / 5) < 1e-12


class _FakeModel:
    """Stands in for the detector: result = the batch's ids (the collection logic is what is under test)."""
    test_cfg = {}
    use_cuda_graph = False

    def eval(self):
        return self

    def forward_test(self, ids=None, **kw):
        return dict(bbox_ids=list(ids), preds=np.asarray(ids, dtype=np.float32)[:, None])


class _Loader(list):
    dataset = None


def _collect_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from edgecape_b200.apis import multi_gpu_test
    # DistributedSampler semantics: rank r sees samples r, r + world, ... (padded by wrapping); batch size 1
    n = 5
    idx = list(range(n)) + [0] * ((-n) % world)
    loader = _Loader([dict(ids=[i]) for i in idx[rank::world]])
    loader.dataset = list(range(n))
    out = multi_gpu_test(_FakeModel(), loader)
    if rank == 0:
        q.put([r["bbox_ids"][0] for r in out])
    else:
        assert out is None
    dist.destroy_process_group()


def test_multi_gpu_test_interleaves_and_truncates_like_the_reference():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_collect_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [0, 1, 2, 3, 4]          # zip(*parts) re-interleaving, padding sample dropped
