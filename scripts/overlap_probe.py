"""How well do the backbone graph of one batch and the head graph of another share the GPU?
Times N replays of (a) the ViT graph alone, (b) the head graph alone, (c) both on two streams concurrently."""
import sys

import torch

sys.path.insert(0, ".")
import edgecape_b200 as E  # noqa: E402
from edgecape_b200.config import default_model_cfg, state_dict_shapes  # noqa: E402
from edgecape_b200.synthetic import make_episode, make_state_dict  # noqa: E402

import os  # noqa: E402
from edgecape_b200 import _lib  # noqa: E402
if os.environ.get("EDGECAPE_GEMM_CTAS"):
    _lib.call("ec_tc_set_cta_limit", int(os.environ["EDGECAPE_GEMM_CTAS"]))
cfg = default_model_cfg("dinov2_vitb14")
model = E.build_model(dict(model=cfg))
model.load_state_dict(make_state_dict(state_dict_shapes(cfg), 0), strict=True)
model = model.cuda().eval()
if os.environ.get("EC_PDL_MODE"):
    model.test_cfg = dict(model.test_cfg, pdl=os.environ["EC_PDL_MODE"])
d = make_episode(batch=16, image_size=256, num_kpts=100, shots=1, seed=1, pin_memory=True)
for _ in range(3):
    model(return_loss=False, **d)
torch.cuda.synchronize()
eng = next(iter(model._graphs.values()))
s0, s1 = eng.slots
N = 20


def timed(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / N


def vit_only():
    with torch.cuda.stream(eng.vit_stream):
        for _ in range(N):
            s0.graph_vit.replay()


def head_only():
    with torch.cuda.stream(eng.head_stream):
        for _ in range(N):
            s1.graph_head.replay()


def both():
    for _ in range(N):
        with torch.cuda.stream(eng.vit_stream):
            s0.graph_vit.replay()
        with torch.cuda.stream(eng.head_stream):
            s1.graph_head.replay()


res = {}
for name, fn in (("vit", vit_only), ("head", head_only), ("both", both), ("vit", vit_only), ("head", head_only), ("both", both)):
    res[name] = min(res.get(name, 1e9), timed(fn))
print(" ".join(f"{k}={v:.3f}" for k, v in res.items()), flush=True)
