"""Where does the end-to-end loop (pinned host tensors in, numpy result dicts out) lose time against resident inputs?
Host-side wall time of forward_test_async (submit) and of PendingResult.result() per batch, and the loop's throughput,
for pipeline depths 2 and 3."""
import sys
import time

import torch

sys.path.insert(0, ".")
import edgecape_b200 as E  # noqa: E402
from edgecape_b200.config import default_model_cfg, state_dict_shapes  # noqa: E402
from edgecape_b200.synthetic import make_episode, make_state_dict  # noqa: E402

cfg = default_model_cfg("dinov2_vitb14")
for depth in (2, 3):
    model = E.build_model(dict(model=cfg))
    model.load_state_dict(make_state_dict(state_dict_shapes(cfg), 0), strict=True)
    model = model.cuda().eval()
    model.test_cfg = dict(model.test_cfg, pipeline_depth=depth)
    host = [make_episode(batch=16, image_size=256, num_kpts=100, shots=1, seed=1234 + i, pin_memory=True) for i in range(4)]
    for _ in range(2):
        for d in host:
            model(return_loss=False, **d)
    torch.cuda.synchronize()
    n = 60
    pend, t_sub, t_res = [], 0.0, 0.0
    t0 = time.perf_counter()
    for i in range(n):
        d = host[i % 4]
        a = time.perf_counter()
        pend.append(model.forward_test_async(**{k: v for k, v in d.items()}))
        b = time.perf_counter()
        t_sub += b - a
        if len(pend) >= depth:
            pend.pop(0).result()
            t_res += time.perf_counter() - b
    while pend:
        pend.pop(0).result()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"depth {depth}: {16 * n / dt:.1f} query img/s, {dt / n * 1e3:.3f} ms per batch; host: submit {t_sub / n * 1e3:.3f} ms, "
          f"result() {t_res / n * 1e3:.3f} ms per batch (waits included)", flush=True)
    # the same with device-resident inputs
    dev = [dict(img_s=[t.cuda() for t in d["img_s"]], img_q=d["img_q"].cuda(), target_s=[t.cuda() for t in d["target_s"]],
                target_weight_s=[t.cuda() for t in d["target_weight_s"]], img_metas=d["img_metas"]) for d in host]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hs = []
    for i in range(n):
        d = dev[i % 4]
        hs.append(model.predict_async(d["img_s"], d["target_s"], d["target_weight_s"], d["img_q"], d["img_metas"]))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"depth {depth}: resident {16 * n / dt:.1f} query img/s, {dt / n * 1e3:.3f} ms per batch", flush=True)
    del model
