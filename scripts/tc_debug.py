"""Stand-alone bring-up of the tcgen05 GEMM (run on the GPU box, each stage under its own timeout)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from edgecape_b200 import ops  # noqa: E402


def check(M, N, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.02
    D = torch.device("cuda")
    a2, b2 = ops.split_f16(x.to(D)), ops.split_f16(w.to(D), 1024.0)
    got = ops.gemm_tc(a2, b2)
    torch.cuda.synchronize()
    want = (x.double() @ w.double().T).float()
    g_ = got.cpu()
    err = (g_ - want).abs().max().item() / want.abs().max().item()
    print(f"tc {M}x{N}x{K}: rel err {err:.3e}  got[0,:4]={g_[0, :4].tolist()} want[0,:4]={want[0, :4].tolist()}", flush=True)
    if err > 1e-4:
        # diagnostics: hi*hi only? row/col permutation?
        hh = (a2.data[:, :a2.Kp].float() @ b2.data[:, :b2.Kp].float().T / 1024.0).cpu()
        print("   vs hi*hi only:", ((g_ - hh).abs().max() / hh.abs().max()).item())
        print("   nonzero frac:", (g_ != 0).float().mean().item(), " nan:", torch.isnan(g_).any().item())
        for r in (0, 1, 8, 32, 64, 127):
            if r < M:
                best = (want - g_[r][None]).abs().sum(1).argmin().item()
                print(f"   got row {r} is closest to want row {best}")
    return err


def bench(M, N, K, iters=20, simt=True):
    D = torch.device("cuda")
    x = torch.randn(M, K, device=D)
    w = torch.randn(N, K, device=D) * 0.02
    a2, b2 = ops.split_f16(x), ops.split_f16(w, 1024.0)
    out = torch.empty(M, N, device=D)
    for _ in range(3):
        ops.gemm_tc(a2, b2, out=out)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        ops.gemm_tc(a2, b2, out=out)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    if not simt:
        print(f"   {M}x{N}x{K}: {ms:.3f} ms = {tf:.1f} algorithmic TFLOP/s", flush=True)
        return
    s.record()
    for _ in range(iters):
        ops.gemm(x, w, out=out)
    e.record()
    torch.cuda.synchronize()
    ms2 = s.elapsed_time(e) / iters
    print(f"bench {M}x{N}x{K}: tc {ms:.3f} ms = {tf:.1f} algorithmic TFLOP/s ({3 * tf:.1f} issued); simt {ms2:.3f} ms "
          f"= {2.0 * M * N * K / ms2 / 1e9:.1f} TFLOP/s", flush=True)


def experiments():
    """Where does the main loop lose time?  (results are wrong while a flag is set)"""
    from edgecape_b200 import _lib
    lib = _lib.load()
    for bn in (128, 256):
        lib.ec_tc_set_tile_n(bn)
        for flags, what in ((0, "baseline"), (4, "no epilogue stores"), (1, "no TMA (operands resident)"),
                            (2, "hi*hi only (1 product, same loads)"), (3, "hi*hi only + no TMA"), (5, "no TMA + no stores"),
                            (7, "hi*hi only, no TMA, no stores")):
            lib.ec_tc_set_debug(flags)
            print(f"[tile {bn}] {what}:")
            bench(10400, 2304, 768, simt=False)
            bench(10400, 768, 3072, simt=False)
    lib.ec_tc_set_debug(0)
    lib.ec_tc_set_tile_n(0)


if __name__ == "__main__":
    stage = sys.argv[1]
    if stage == "pair":
        from edgecape_b200 import _lib
        _lib.load().ec_tc_set_tile_n(512)
        check(256, 256, 64)
        check(256, 256, 256)
        check(512, 768, 768)
        check(1300, 768, 768)
        check(200, 96, 100)
        check(650, 3072, 768)
        check(10400, 2304, 768)
        sys.exit(0)
    if stage == "modes":
        # every tile mode on the ViT and head shapes: input to the mode heuristic of ec_gemm_f16x3
        from edgecape_b200 import _lib
        shapes = [(10400, 2304, 768), (10400, 768, 768), (10400, 3072, 768), (10400, 768, 3072), (5184, 256, 768),
                  (5184, 1024, 512), (5184, 512, 512), (5184, 256, 512), (1600, 512, 512), (1600, 768, 256),
                  (1600, 256, 512), (1600, 2048, 256), (1600, 256, 2048), (1600, 256, 256), (6784, 768, 256),
                  (6784, 2048, 256), (6784, 256, 2048), (6784, 256, 256), (1600, 512, 1024)]
        for shp in shapes:
            print("shape", shp)
            for bn in (128, 256, 512):
                if bn == 512 and shp[1] < 256:
                    continue
                _lib.load().ec_tc_set_tile_n(bn)
                print(f"  mode {bn}:", end="")
                bench(*shp, iters=30, simt=False)
        sys.exit(0)
    if stage == "pairbench":
        from edgecape_b200 import _lib
        for bn in (512, 128):
            _lib.load().ec_tc_set_tile_n(bn)
            print("tile mode", bn)
            for shp in [(10400, 2304, 768), (10400, 768, 768), (10400, 3072, 768), (10400, 768, 3072), (5184, 256, 768)]:
                bench(*shp, simt=False)
        sys.exit(0)
    if stage == "epi":
        # the four ViT-B GEMMs with their real epilogues (M = 32 images x 325 tokens)
        D = torch.device("cuda")
        M, C = 10400, 768
        x = torch.randn(M, C, device=D)
        h = torch.randn(M, 4 * C, device=D)
        t = torch.randn(M, C, device=D)
        g = torch.rand(C, device=D)
        ws = {n: ops.split_f16(torch.randn(o, i, device=D) * 0.02, 1024.0) for n, (o, i) in
              dict(qkv=(3 * C, C), proj=(C, C), fc1=(4 * C, C), fc2=(C, 4 * C)).items()}
        bs = {n: torch.randn(w.rows, device=D) for n, w in ws.items()}
        x2, h2 = ops.split_f16(x), ops.split_f16(h)
        qkv = torch.empty(M, 3 * C, device=D)
        cases = dict(
            qkv=lambda: ops.gemm_tc(x2, ws["qkv"], out=qkv, bias=bs["qkv"]),
            proj=lambda: ops.gemm_tc(x2, ws["proj"], out=t, bias=bs["proj"], colscale=g, residual=t),
            fc1=lambda: ops.gemm_tc(x2, ws["fc1"], bias=bs["fc1"], act=ops.ACT_GELU, split_out=True, fp32_out=False),
            fc2=lambda: ops.gemm_tc(h2, ws["fc2"], out=t, bias=bs["fc2"], colscale=g, residual=t))
        from edgecape_b200 import _lib
        for mode in (512, 128):
            _lib.load().ec_tc_set_tile_n(mode)
            tot = 0.0
            for n, fn in cases.items():
                for _ in range(3):
                    fn()
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record()
                for _ in range(20):
                    fn()
                e_.record()
                torch.cuda.synchronize()
                ms = s_.elapsed_time(e_) / 20
                tot += ms
                K = 4 * C if n == "fc2" else C
                print(f"[mode {mode}] {n}: {ms:.3f} ms = {2.0 * M * ws[n].rows * K / ms / 1e9:.1f} algorithmic TFLOP/s", flush=True)
            print(f"[mode {mode}] one ViT-B layer of GEMMs: {tot:.3f} ms")
        sys.exit(0)
    if stage == "exp":
        experiments()
        sys.exit(0)
    if stage == "tiny":
        check(128, 128, 64)
    elif stage == "k2":
        check(128, 128, 256)
    elif stage == "multi":
        from edgecape_b200 import _lib
        for bn in (128, 256):
            _lib.load().ec_tc_set_tile_n(bn)
            print("tile width", bn)
            check(512, 384, 768)
            check(1300, 768, 768)
            check(200, 96, 100)
            check(650, 3072, 768)
    elif stage == "bench":
        from edgecape_b200 import _lib
        for bn in (128, 256):
            _lib.load().ec_tc_set_tile_n(bn)
            print("tile width", bn)
            for shp in [(10400, 2304, 768), (10400, 768, 768), (10400, 3072, 768), (10400, 768, 3072)]:
                bench(*shp)
        _lib.load().ec_tc_set_tile_n(0)
        print("tile width: heuristic")
        for shp in [(10400, 2304, 768), (10400, 768, 768), (10400, 3072, 768), (10400, 768, 3072), (5184, 256, 768),
                    (1600, 256, 256)]:
            bench(*shp)
