"""Why is the end-to-end loop ~4 % below the resident loop?  Same pipelined test loop (forward_test_async / result(), depth 2),
with the inputs (a) all on the host (pinned), (b) images on the host and heat-map targets on the device, (c) all on the
device -- plus the raw H2D bandwidth of the box for the step's 51 MB."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import edgecape_b200 as E  # noqa: E402
from edgecape_b200.config import default_model_cfg, state_dict_shapes  # noqa: E402
from edgecape_b200.synthetic import make_episode, make_state_dict  # noqa: E402

cfg = default_model_cfg("dinov2_vitb14")
model = E.build_model(dict(model=cfg))
model.load_state_dict(make_state_dict(state_dict_shapes(cfg), 0), strict=True)
model = model.cuda().eval()
host = [make_episode(batch=16, image_size=256, num_kpts=100, shots=1, seed=1234 + i, pin_memory=True) for i in range(4)]

# raw H2D bandwidth
src = torch.empty(51 * 2 ** 20, dtype=torch.uint8).pin_memory()
dst = torch.empty_like(src, device="cuda")
for _ in range(3):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    dst.copy_(src, non_blocking=True)
b.record()
torch.cuda.synchronize()
print(f"H2D pinned: {51 * 2 ** 20 * 10 / (a.elapsed_time(b) * 1e-3) / 1e9:.1f} GB/s ({a.elapsed_time(b) / 10:.3f} ms per 51 MiB)", flush=True)


def variant(kind):
    out = []
    for d in host:
        e = dict(d)
        if kind in ("tgt_dev", "all_dev"):
            e["target_s"] = [t.cuda() for t in d["target_s"]]
            e["target_weight_s"] = [t.cuda() for t in d["target_weight_s"]]
        if kind == "all_dev":
            e["img_s"] = [t.cuda() for t in d["img_s"]]
            e["img_q"] = d["img_q"].cuda()
        out.append(e)
    return out


def loop(batches, n, depth=2):
    pend = []
    for i in range(n):
        pend.append(model.forward_test_async(**batches[i % 4]))
        if len(pend) >= depth:
            pend.pop(0).result()
    while pend:
        pend.pop(0).result()


for rep in range(2):
    for kind in ("all_host", "tgt_dev", "all_dev"):
        bt = variant(kind)
        loop(bt, 8)
        torch.cuda.synchronize()
        n = 60
        t0 = time.perf_counter()
        loop(bt, n)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{kind:9s}: {16 * n / dt:7.1f} query img/s, {dt / n * 1e3:.3f} ms per batch", flush=True)
    # resident, host running ahead (what bench.py's `value` times)
    bt = variant("all_dev")
    t0 = time.perf_counter()
    for i in range(60):
        d = bt[i % 4]
        model.predict_async(d["img_s"], d["target_s"], d["target_weight_s"], d["img_q"], d["img_metas"])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"resident : {16 * 60 / dt:7.1f} query img/s, {dt / 60 * 1e3:.3f} ms per batch", flush=True)
