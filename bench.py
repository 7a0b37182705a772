#!/usr/bin/env python
"""bench.py -- query images / second of the EdgeCape inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 1-shot, 256x256 (-> 18x18 patches), DINOv2 ViT-B/14,
100 keypoints, batch 16 per GPU, synthetic episodes and deterministic random-init weights
(edgecape_b200/synthetic.py).  One step = one batch through the whole path: (1+shots) ViT forwards
per query + head (skeleton predictor, encoder, proposals, graph decoder) + on-device PCK counters.
Queries shard over GPUs with no data-path collective (weak scaling: per-GPU batch fixed); one NCCL
all-reduce of the fp64 PCK counters closes the run.

One JSON line is printed by rank 0; see the task contract for the keys.  `value` is measured with
inputs resident in HBM (`model.predict_async`, the library's two-deep pipeline over consecutive
steps: every one of the K steps starts and completes inside the timed region); `e2e` goes through
the reference-facing test loop `edgecape_b200.apis.iter_results` (= `single_gpu_test`, one
`model.forward_test_async(**data)` per batch, result dicts consumed in order) with pinned host
tensors, every H2D and D2H copy inside the timed region.  `roofline` covers the
dominant kernel (the ViT/head GEMM), timed live with CUDA events on the launching stream.
`cpu_baseline` / `--impl reference` time the CPU restatement of the reference (oracle/, pinned
against the unmodified reference by tests/golden) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "query images/sec (1-shot, 256x256, ViT-B/14, 100 kpts)"
WORKLOAD = "configs[1]: 1-shot synthetic 256x256, DINOv2-B/14, 100-kpt random skeleton, batch 16 per GPU"

# --config presets: BASELINE.json configs[1..4] as per-GPU slices (the driver's default run is c2 = configs[1]; the
# others are parity-test cases whose per-GPU throughput is recorded under profiles/, not bench lines of the round)
PRESETS = {
    "c2": dict(backbone="dinov2_vitb14", image_size=256, kpts=100, shots=1, batch=16, skeleton="tree+extra",
               workload=WORKLOAD, metric=METRIC),
    "c3": dict(backbone="dinov2_vitb14", image_size=256, kpts=100, shots=1, batch=16, skeleton="tree+extra",
               workload="configs[2]: 1-shot synthetic 256x256, DINOv2-B/14, random 100-kpt skeleton, batch 128 sharded "
                        "over 8 GPUs = 16 per GPU"),
    "c4": dict(backbone="dinov2_vitb14", image_size=256, kpts=100, shots=5, batch=8, skeleton="tree+extra",
               workload="configs[3]: 5-shot synthetic 256x256, DINOv2-B/14, 100 kpts, batch 64 over 8 GPUs = 8 per GPU"),
    "c5": dict(backbone="dinov2_vitl14", image_size=384, kpts=200, shots=1, batch=4, skeleton="full",
               workload="configs[4]: dense-graph stress, 1-shot synthetic 384x384, DINOv2-L/14, 200-kpt fully-connected "
                        "skeleton (19 900 edges), batch 32 over 8 GPUs = 4 per GPU"),
}


def describe(args):
    """(metric, workload) strings: BASELINE.json's wording for the default arguments (configs[1]), the preset's
    wording for --config, an explicit description otherwise so that a non-default run can never be mistaken for the
    headline."""
    generic = (f"query images/sec ({args.shots}-shot, {args.image_size}x{args.image_size}, {args.backbone}, "
               f"{args.kpts} kpts)")
    if args.config:
        pr = PRESETS[args.config]
        return pr.get("metric", generic), pr["workload"]
    default = (args.backbone == "dinov2_vitb14" and args.image_size == 256 and args.kpts == 100 and args.shots == 1 and
               args.batch == 16 and args.skeleton == "tree+extra")
    if default:
        return METRIC, WORKLOAD
    return (generic, f"NON-DEFAULT: {args.shots}-shot synthetic {args.image_size}x{args.image_size}, {args.backbone}, "
                     f"{args.kpts}-kpt {args.skeleton} skeleton, batch {args.batch} per GPU")


def workload_config(args, world):
    """The `config` object of the JSON line -- identical for both arms (`--impl ours` / `--impl reference`)."""
    return {"workload": describe(args)[1], "per_gpu_batch": args.batch, "global_batch": args.batch * world,
            "image_size": args.image_size, "keypoints": args.kpts, "shots": args.shots, "backbone": args.backbone,
            "skeleton": args.skeleton}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default=None, choices=sorted(PRESETS),
                    help="BASELINE.json configs[1..4] as per-GPU slices (overrides the shape arguments)")
    ap.add_argument("--batch", type=int, default=16, help="queries per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--backbone", default="dinov2_vitb14")
    ap.add_argument("--dedup-demo", action="store_true",
                    help="also time episodes whose queries share their support sample, with and without de-duplication")
    ap.add_argument("--image-size", type=int, default=256)
    ap.add_argument("--kpts", type=int, default=100)
    ap.add_argument("--shots", type=int, default=1)
    ap.add_argument("--skeleton", default="tree+extra", choices=["tree+extra", "chain", "full"])
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="queries per CPU-baseline step (0 = min(batch, 16), BASELINE.md section 3)")
    ap.add_argument("--sustained-seconds", type=float, default=10.0,
                    help="length of the extra sustained-throughput pass (clock sampler on); 0 disables it")
    ap.add_argument("--depth", type=int, default=0, help="batches in flight in the library's pipeline (0 = its default, 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.config:
        for k, v in PRESETS[args.config].items():
            if k not in ("workload", "metric"):
                setattr(args, k, v)
    if args.cpu_sample <= 0:
        args.cpu_sample = min(args.batch, 16)
    return args


def q_per_step_local(B, world):
    return B * world


def model_cfg(backbone):
    from edgecape_b200.config import default_model_cfg
    return default_model_cfg(backbone)


def flops_per_query(cfg, image_size, K, shots):
    """Algorithmic FLOPs (multiply-add = 2) of one query, SURVEY.md section 8d."""
    from edgecape_b200.config import vit_config
    v = vit_config(cfg["pretrained"])
    C, L, P = v["embed_dim"], v["depth"], v["patch_size"]
    S = (image_size // P) ** 2
    N = S + 1
    vit = L * (2 * N * 12 * C * C + 4 * N * N * C) + 2 * S * 3 * P * P * C
    return (1 + shots) * vit


# --------------------------------------------------------------------------- CPU baseline
def cpu_episode(args):
    """The CPU arm's batch: min(B, 16) queries of the same workload; with the default batch
    it is exactly rank 0's first GPU batch (same generator arguments)."""
    from edgecape_b200.synthetic import make_episode
    return make_episode(batch=max(1, args.cpu_sample), image_size=args.image_size, num_kpts=args.kpts,
                        shots=args.shots, seed=1234, skeleton=args.skeleton)


def cpu_reference_rate(args, steps, warmup):
    """Times oracle.edgecape_oracle.detector_forward_test (CPU restatement of the reference's
    forward_test, pinned to the unmodified reference by tests/golden) on the host cores.
    Returns (cpu_baseline object, ms per step, the oracle's outputs on that batch)."""
    from edgecape_b200.config import state_dict_shapes
    from edgecape_b200.synthetic import make_state_dict
    from oracle import edgecape_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = model_cfg(args.backbone)
    sd = make_state_dict(state_dict_shapes(cfg), 0)
    data = cpu_episode(args)
    b = data["img_q"].shape[0]
    times, out = [], None
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = edgecape_oracle.detector_forward_test(sd, cfg, data, torch.float32)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    return dict(value=b / (ms / 1e3), unit="query images/s", cores=cores, kind="port",
                sample=f"{b} queries/step x {steps} steps (+{warmup} warm-up) of the same workload, fp32, "
                       f"torch CPU {cores} threads, oracle/edgecape_oracle.py"), ms, out


def parity_against_oracle(model, args, want):
    """The GPU path against the CPU oracle on the CPU arm's batch (= the first timed GPU batch at the default
    arguments): relative-to-max error of the result dict through the public call (CUDA-graph engine, pinned copies),
    and of the heat-map / proposals / per-layer coordinates through the eager path, whose arg-max keypoint indices
    must equal the oracle's."""
    data = cpu_episode(args)
    res = model(return_loss=False, **data)
    _, inter = model.predict(data["img_s"], data["target_s"], data["target_weight_s"], data["img_q"], data["img_metas"],
                             return_intermediates=True)
    errs = {}

    def rel(a, w):
        a = a.detach().float().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
        w = w.detach().float().cpu().numpy() if torch.is_tensor(w) else np.asarray(w)
        return float(np.abs(a.astype(np.float64) - w).max() / (np.abs(w).max() + 1e-12))

    for k in ("preds", "points", "skeleton"):
        errs[k] = rel(res[k], want[k])
    for k in ("similarity_map", "initial_proposals", "output", "adj"):
        if k in inter and inter[k] is not None and k in want:
            errs[k] = rel(inter[k], want[k])
    am = inter["argmax"].detach().cpu().numpy().astype(np.int64)
    wm = want["argmax"].detach().cpu().numpy().astype(np.int64)
    return {"max_rel_err": max(errs.values()), "argmax_equal": bool(np.array_equal(am, wm)),
            "argmax_compared": int(wm.size), "rel_err": errs, "queries": int(data["img_q"].shape[0]),
            "bar": "1e-3 relative to the tensor's max magnitude, arg-max keypoint indices bit-exact",
            "against": "oracle/edgecape_oracle.py fp32 on the same batch (pinned to the unmodified reference by tests/golden)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, min(args.steps, 100)), max(1, min(args.warmup, 10))
    cb, ms, _ = cpu_reference_rate(args, steps, warmup)
    line = {
        "impl": "reference", "metric": describe(args)[0], "value": cb["value"], "unit": "query images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "note": "reference's PyTorch-CPU forward_test restated in oracle/ (the reference itself cannot travel to the "
                "GPU box: mmcv/mmpose/hub DINOv2 are absent); each step is one batch of min(per_gpu_batch, 16) queries "
                "of the workload on all host threads",
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "query images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


# -------------------------------------------------------------------------------- ours
class GemmTimer:
    """CUDA-event timing of every ec_gemm launch issued while active (roofline of the dominant kernel)."""

    def __init__(self):
        self.records = []      # (flops, start, end)
        self.active = False

    def install(self):
        from edgecape_b200 import _lib, ops
        orig = _lib.call
        timer = self

        def call(name, *a):
            if timer.active and name in ("ec_gemm", "ec_gemm_f16x3", "ec_gemm_f16f8"):
                M, N, K = a[3], a[4], a[5]
                batch = a[10] if name == "ec_gemm" else 1
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                orig(name, *a)
                e.record()
                timer.records.append((2.0 * M * N * K * batch, (name, M, N, K, batch), s, e))
            else:
                orig(name, *a)

        _lib.call = call
        ops._lib.call = call

    def summary(self):
        big = [(f, shp, s.elapsed_time(e)) for f, shp, s, e in self.records]
        if not big:
            return None
        # dominant kernel = the large ViT-shaped launches (>= 1 GFLOP each)
        dom = [r for r in big if r[0] >= 1e9] or big
        fl = sum(r[0] for r in dom)
        ms = sum(r[2] for r in dom)
        tc = sum(1 for r in dom if r[1][0] in ("ec_gemm_f16x3", "ec_gemm_f16f8"))
        f8 = sum(1 for r in dom if r[1][0] == "ec_gemm_f16f8")
        # tensor-pipe time issued, in units of one fp16 product: 3 for three fp16 products, 2 for fp16 + two e4m3
        # cross terms (an e4m3 UMMA covers twice the contraction depth per clock)
        units = sum(r[0] * (2.0 if r[1][0] == "ec_gemm_f16f8" else 3.0) for r in dom if r[1][0] != "ec_gemm")
        return dict(flops=fl, ms=ms, launches=len(dom), all_ms=sum(r[2] for r in big), all_launches=len(big),
                    tensor_core_launches=tc, f8_launches=f8, issued_units=units)


def north_star_kernels(peaks):
    """The two kernels the north-star names, timed alone at its shapes (CUDA-graph replay of 20 launches between
    CUDA events, warm; inputs of one launch are far smaller than L2, so these are L2-resident figures):
    the skeleton-GCN feed-forward at batch 64 against the HBM roofline (SURVEY 8d bytes) and the ViT patch attention
    of one bench step against the tensor-pipe roofline (algorithmic 4 N^2 D per head; 3 fp16 products issued per
    algorithmic product)."""
    import torch
    from edgecape_b200 import ops
    dev = torch.device("cuda", torch.cuda.current_device())
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    tf = float(peaks.get("bf16_tflops", 1650.0))          # a kernel timed alone: the burst figure

    def graph_time(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) * 1e3 / iters          # us per launch

    out = {}
    B, K, d, dff = 64, 100, 256, 384
    x = torch.randn(B, K, d, device=dev)
    adj = ops.soft_normalize_adj(torch.rand(B, K, K, device=dev), torch.zeros(B, K, dtype=torch.uint8, device=dev))
    Wp = ops.gcn_pack_weights(torch.randn(2 * dff, d, device=dev) * d ** -0.5, torch.randn(2 * dff, device=dev) * 0.1)
    y = torch.empty(B, K, dff, device=dev)
    us = graph_time(lambda: ops.gcn(x, adj, Wp, out=y))
    nbytes = B * K * d * 4 + B * K * K * 4 + B * K + (2 * dff * d + 2 * dff) * 4 + B * K * dff * 4
    out["gcn"] = {"shape": f"B={B} K={K} d={d} dff={dff}", "us": us, "bound": "hbm", "achieved": nbytes / us / 1e3,
                  "peak": hbm, "unit": "GB/s", "frac": nbytes / us / 1e3 / hbm,
                  "kernel": ("ec::gf2::gcn_fused2_kernel (project first, persistent)" if ops.GCN_FUSED >= 2 else "ec::gf::gcn_fused_kernel")
                  if ops.gcn_fused_ok(B, K, d, dff) else "gcn_aggregate_split + gemm_f16x3",
                  "note": "3.0 algorithmic GFLOP (6.4 issued: fp16 hi.hi + two e4m3 cross terms for X W^T, three fp16 products for "
                          "A1 T1): the tensor time alone is ~4.5 us, the byte roofline 3.0 us; at batch 64 the launch is one wave, "
                          "i.e. the critical path of one CTA"}
    Bi, H, N, D = 32, 12, 325, 64
    if ops.attention_split_ok(D, N):
        qkv2 = ops.split_f16(torch.randn(Bi * N, 3 * H * D, device=dev))
        us = graph_time(lambda: ops.attention_packed_split(qkv2, Bi, N, H))
        fl = 4.0 * N * N * D * H * Bi
        out["vit_attention"] = {"shape": f"{Bi} images x {H} heads x {N} tokens x {D}", "us": us, "bound": "tensor",
                                "achieved": fl / us / 1e6, "peak": tf, "unit": "TFLOP/s", "frac": fl / us / 1e6 / tf,
                                "issued_frac": 3 * fl / us / 1e6 / tf,
                                "kernel": "ec::atc::attention_tc_ps_kernel (persistent, software-pipelined over tiles)"}
    return out


def run_ours(args):
    import torch.distributed as dist
    from edgecape_b200 import _lib, ops, build_model
    from edgecape_b200.config import state_dict_shapes
    from edgecape_b200.synthetic import make_episode, make_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    if os.environ.get("EDGECAPE_GEMM_CTAS"):
        _lib.call("ec_tc_set_cta_limit", int(os.environ["EDGECAPE_GEMM_CTAS"]))
    cfg = model_cfg(args.backbone)
    model = build_model(dict(model=cfg))
    model.load_state_dict(make_state_dict(state_dict_shapes(cfg), 0), strict=True)
    model = model.cuda().eval()
    model.use_cuda_graph = not args.no_graph
    if args.depth > 0:
        model.test_cfg = dict(model.test_cfg, pipeline_depth=args.depth)

    B, R, K = args.batch, args.image_size, args.kpts
    NB = 4   # distinct input batches rotated through the timed region
    host = [make_episode(batch=B, image_size=R, num_kpts=K, shots=args.shots, seed=1234 + 97 * rank + i,
                         pin_memory=True, skeleton=args.skeleton) for i in range(NB)]
    devb = []
    for d in host:
        devb.append(dict(img_s=[t.to(dev) for t in d["img_s"]], img_q=d["img_q"].to(dev),
                         target_s=[t.to(dev) for t in d["target_s"]],
                         target_weight_s=[t.to(dev) for t in d["target_weight_s"]], img_metas=d["img_metas"]))
    gt = [torch.rand(B, K, 2, device=dev) for _ in range(NB)]
    valid = [(d["target_weight_s"][0].reshape(B, K) > 0).to(torch.uint8).contiguous() for d in devb]
    norm = torch.full((B, 2), 1.0, device=dev)
    thr = torch.tensor([0.05, 0.1, 0.15, 0.2, 0.25], device=dev)
    counters = torch.zeros(8, dtype=torch.float64, device=dev)

    from edgecape_b200.apis import iter_results
    from edgecape_b200.parallel import allreduce_counters, summarize_pck

    def run_resident(start, n):
        """n steps on device-resident inputs.  Graph mode: the library's depth-2 pipeline (backbone of step i+1 beside
        the head of step i); the PCK counters of a step are accumulated on the stream its outputs are ordered on."""
        for i in range(start, start + n):
            d = devb[i % NB]
            if model.use_cuda_graph:
                h = model.predict_async(d["img_s"], d["target_s"], d["target_weight_s"], d["img_q"], d["img_metas"])
                with torch.cuda.stream(h.stream):
                    ops.pck_accumulate_(counters, h.out[0][-1], gt[i % NB], valid[i % NB], norm, thr)
            else:
                out = model.predict(d["img_s"], d["target_s"], d["target_weight_s"], d["img_q"], d["img_metas"])[0]
                ops.pck_accumulate_(counters, out[-1], gt[i % NB], valid[i % NB], norm, thr)

    def run_e2e(start, n):
        """n steps through the public test loop (edgecape_b200.apis, the drop-in for the reference's single_gpu_test):
        pinned HOST tensors in, result dicts (numpy) out, every H2D / D2H copy inside the timed region."""
        last = None
        for res in iter_results(model, (host[i % NB] for i in range(start, start + n))):
            last = res["preds"]
        return last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, timer=None, reduce=False):
        fn(0, warmup)
        barrier()
        if reduce:
            counters.zero_()
            torch.cuda.synchronize()
        if timer is not None:
            timer.active = True
        l0 = _lib.launch_count()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn(warmup, steps)
        if reduce:
            # the single collective of the path closes the last step INSIDE the timed region: the fp64 PCK
            # counters of every rank, summed over NVLink by NCCL, ordered after the last step's counter kernel
            if model.use_cuda_graph:
                for g in model._graphs.values():
                    torch.cuda.current_stream().wait_stream(g.head_stream)
            allreduce_counters(counters)
        torch.cuda.synchronize()           # the pipeline's streams: every step has completed before the end event
        e.record()
        if timer is not None:
            timer.active = False
        barrier()
        ms = max(s.elapsed_time(e), 0.0)
        launches = _lib.launch_count() - l0
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    gt_timer = GemmTimer()
    gt_timer.install()
    W = max(3, args.warmup)
    with ClockSampler(local) as clocks:
        # (1) headline: the CUDA-graph replay path (what model(...) runs by default)
        ms_total, _ = timed(run_resident, args.steps, W, reduce=True)
        pck_counters = counters.clone()
        ms_e2e, _ = timed(run_e2e, args.steps, W)
        # (2) the same K steps launched eagerly, with a CUDA-event pair around every GEMM launch: per-kernel
        #     durations for the roofline, and the count of kernels one step launches (a graph replays them)
        model.use_cuda_graph = False
        ms_eager, launches = timed(run_resident, args.steps, 2, gt_timer)
        model.use_cuda_graph = not args.no_graph
        roof = gt_timer.summary()
    # informational (NOT the headline: SURVEY 8d counts (1+shots) backbone passes per query): the same loop on episodes
    # whose 16 queries share one support sample, as MP-100 test episodes do, with support de-duplication on
    dedup = None
    if not args.no_graph and args.dedup_demo:
        shared = [make_episode(batch=B, image_size=R, num_kpts=K, shots=args.shots, seed=4321 + 97 * rank + i,
                               pin_memory=True, shared_support=B) for i in range(NB)]
        def run_shared(start, n):
            for res in iter_results(model, (shared[i % NB] for i in range(start, start + n))):
                pass
        ms_plain, _ = timed(run_shared, args.steps, W)
        model.test_cfg = dict(model.test_cfg, dedup_supports=True)
        ms_dedup, _ = timed(run_shared, args.steps, W)
        model.test_cfg = dict(model.test_cfg, dedup_supports=False)
        dedup = {"episodes": f"{B} queries share 1 support sample per batch",
                 "e2e_value_reference_batching": q_per_step_local(B, world) * args.steps / (ms_plain / 1e3),
                 "e2e_value_dedup_supports": q_per_step_local(B, world) * args.steps / (ms_dedup / 1e3),
                 "unit": "query images/s"}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # the two kernels the north-star names, alone at its shapes -- BEFORE the sustained pass: after ten seconds under the
    # power cap the SM clock sits ~20 % lower for a while, and a kernel timed then is not comparable with the burst peaks
    nsk = north_star_kernels(peaks) if (rank == 0 and world == 1) else None
    # does the headline survive the power cap?  >= --sustained-seconds of back-to-back resident steps with the clock
    # sampler running (a real evaluation is thousands of steps; the K-step headline region is ~0.1 s)
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(args.steps, int(np.ceil(args.sustained_seconds * 1e3 / (ms_total / args.steps))))
        with ClockSampler(local) as sus_clocks:
            ms_sus, _ = timed(run_resident, n_sus, 2)
        sc = sus_clocks.summary()
        sustained = {"seconds": ms_sus / 1e3, "steps": n_sus, "value": B * world * n_sus / (ms_sus / 1e3),
                     "unit": "query images/s", "ms_per_step": ms_sus / n_sus, "clocks": sc}
    torch.cuda.synchronize()

    # denominator: the per-launch GEMM timings come from a region of K eager steps -- well under a second, SM clocks at
    # boost -- so the BURST figure is the honest peak there; the sustained figure is reported beside it
    burst_region = ms_eager < 1000.0
    peak_sus = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_burst = float(peaks.get("bf16_tflops", 1590.0))
    peak_tf = peak_burst if burst_region else peak_sus
    which = "bf16_tflops (burst: timed region < 1 s)" if burst_region else "bf16_tflops_sustained (timed region >= 1 s)"
    peak_src = (f"MEASURED_PEAKS.json {which} (of measured)" if peaks else f"fallback {peak_tf / 1e3:.2f} PFLOP/s, {which} (of fallback)")
    q_per_step = B * world
    value = q_per_step * args.steps / (ms_total / 1e3)
    e2e_value = q_per_step * args.steps / (ms_e2e / 1e3)
    d0 = host[0]
    h2d = sum(t.numel() * 4 for t in [d0["img_q"]] + d0["img_s"] + d0["target_s"] + d0["target_weight_s"])
    L = cfg["keypoint_head"]["num_decoder_layer"]
    d2h = (B * K * 2 + (1 + L) * B * K * 2 + 2 * K * K) * 4
    line = {
        "metric": describe(args)[0], "value": value, "unit": "query images/s", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "l2": "no explicit flush: weights (0.41 GB) + per-step activations exceed the 126 MB L2 and "
              f"inputs rotate over {NB} distinct batches",
        "e2e": {"value": e2e_value, "unit": "query images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps,
                "api": "edgecape_b200.apis.iter_results (the single_gpu_test loop): model.forward_test_async per batch, "
                       "pinned host tensors in, numpy result dicts out, up to 2 batches in flight"},
        "gpu_launches": int(launches),
        "launch_mode": "eager" if args.no_graph else ("cuda-graph replay of the same kernels (gpu_launches counted on the eager "
                                                      "pass), consecutive steps software-pipelined 2 deep: backbone of step "
                                                      "i+1 beside the head of step i"),
        "eager_ms_per_step": ms_eager / args.steps,
        "clocks": clocks.summary(),
        "collective": "one all_reduce(SUM) of the fp64 PCK counters, inside the timed region after the last step"
                      + (" (NCCL)" if world > 1 else " (no-op at world size 1)"),
        "pck_counters": [float(x) for x in pck_counters.cpu().tolist()[:6]],
        "pck_vs_random_gt": summarize_pck(pck_counters[:6]),
        "algorithmic_gflop_per_query": flops_per_query(cfg, R, K, args.shots) / 1e9,
    }
    if sustained is not None:
        line["sustained"] = sustained
    if dedup is not None:
        line["support_dedup_demo"] = dedup
    traffic = None
    try:
        traffic = json.load(open(os.path.join(REPO, "profiles", "roofline_traffic.json")))["traffic_bytes_per_launch"]
    except Exception:
        pass
    if roof:
        ach = roof["flops"] / (roof["ms"] / 1e3) / 1e12
        line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                            "frac": ach / peak_tf, "traffic": traffic,
                            "kernel": ("ec::tc::gemm_f16x3_kernel (tcgen05, fp32 TMEM accumulate, fp32-grade split operands: "
                                       f"{roof['f8_launches']} of the {roof['launches']} timed launches run a_hi.b_hi on "
                                       "kind::f16 + both cross terms on kind::f8f6f4 e4m3 = 2 units of tensor time per "
                                       "algorithmic product, the others three kind::f16 products = 3 units)")
                            if roof["tensor_core_launches"] else "ec::gemm_simt_kernel<128,128> (fp32 FFMA)",
                            "issued_frac": (roof["issued_units"] / (roof["ms"] / 1e3) / 1e12 / peak_tf)
                            if roof["tensor_core_launches"] else None,
                            "frac_of_sustained_peak": ach / peak_sus, "frac_of_burst_peak": ach / peak_burst,
                            "launches_timed": roof["launches"], "kernel_ms_per_step": roof["ms"] / args.steps,
                            "all_gemm_ms_per_step": roof["all_ms"] / args.steps, "peak_source": peak_src}
    if nsk is not None:
        line["north_star_kernels"] = nsk
    # range events of the e4m3 / fp16 planes over everything this process ran (device-side counters of every F16F8
    # producer): 0 / 0 on the synthetic weights; a real checkpoint with outlier activations would show up here
    try:
        ev = ops.overflow_count()
        line["f16f8_range_events"] = {"beyond_e4m3_448": ev[0], "beyond_fp16_65504": ev[1]}
    except Exception as exc:           # (a diagnostic must not cost the bench line)
        line["f16f8_range_events"] = {"error": str(exc)}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            cb, _, want = cpu_reference_rate(args, 3, 1)
            line["cpu_baseline"] = cb
            line["parity"] = parity_against_oracle(model, args, want)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
