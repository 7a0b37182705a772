"""Stand-alone bring-up of the tcgen05 attention kernel (run on the GPU box under a timeout per stage)."""
import sys

import torch

sys.path.insert(0, ".")
from edgecape_b200 import ops  # noqa: E402


def check(B, H, Lq, Lk, seed=0):
    g = torch.Generator().manual_seed(seed)
    E = H * 64
    q, k, v = (torch.randn(B, L, E, generator=g) for L in (Lq, Lk, Lk))
    qh = q.double().view(B, Lq, H, 64).transpose(1, 2) / 8.0
    kh = k.double().view(B, Lk, H, 64).transpose(1, 2)
    vh = v.double().view(B, Lk, H, 64).transpose(1, 2)
    want = ((qh @ kh.transpose(-1, -2)).softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, E).float()
    D = torch.device("cuda")
    ops.TENSOR_CORES, ops.ATTENTION_TC = True, True
    got = ops.attention(q.to(D), k.to(D), v.to(D), H).cpu()
    torch.cuda.synchronize()
    err = (got - want).abs().max().item() / want.abs().max().item()
    print(f"attn_tc B{B} H{H} Lq{Lq} Lk{Lk}: rel err {err:.3e} nan={torch.isnan(got).any().item()} "
          f"got[0,0,:3]={got[0, 0, :3].tolist()} want={want[0, 0, :3].tolist()}", flush=True)
    if err > 1e-3:
        # is it plain averaging (softmax broken) or permuted rows?
        mean_v = v.view(B, Lk, H, 64).mean(1)[0].reshape(-1)[:3]
        print("   mean of V[0,:,0:3] =", mean_v.tolist())
        for r in (0, 1, 8, 31, 32, 64, 127):
            if r < Lq:
                best = (want[0] - got[0, r][None]).abs().sum(1).argmin().item()
                print(f"   got row {r} closest to want row {best}")
    return err


def bench(B, H, L, iters=20):
    D = torch.device("cuda")
    C = H * 64
    qkv = torch.randn(B, L, 3 * C, device=D)
    out = torch.empty(B, L, C, device=D)
    for tc in (True, False):
        ops.ATTENTION_TC = tc
        for _ in range(3):
            ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H, out=out)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H, out=out)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / iters
        fl = 4.0 * B * H * L * L * 64
        print(f"bench attention B{B} H{H} L{L} {'tcgen05' if tc else 'simt'}: {ms:.3f} ms = {fl / ms / 1e9:.1f} algorithmic TFLOP/s", flush=True)


def check_tma(B, N, H, seed=0):
    g = torch.Generator().manual_seed(seed)
    C = H * 64
    qkv = torch.randn(B * N, 3 * C, generator=g)
    q, k, v = (qkv[:, i * C:(i + 1) * C].double().view(B, N, H, 64).transpose(1, 2) for i in range(3))
    want = (((q / 8.0) @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(B, N, C).float()
    D = torch.device("cuda")
    got = ops.attention_packed_split(ops.split_f16(qkv.to(D)), B, N, H, split="no").cpu()
    torch.cuda.synchronize()
    err = (got - want).abs().max().item() / want.abs().max().item()
    print(f"attn_tma B{B} N{N} H{H}: rel err {err:.3e} nan={torch.isnan(got).any().item()} got={got[0, 0, :3].tolist()} "
          f"want={want[0, 0, :3].tolist()}", flush=True)
    if err > 1e-3:
        mean_v = v.float().mean(2)[0, 0, :3]
        print("   mean of V =", mean_v.tolist(), " (a softmax-less average would give this)")
        for r in (0, 1, 8, 63, 64, 127):
            if r < N:
                best = (want[0] - got[0, r][None]).abs().sum(1).argmin().item()
                print(f"   got row {r} closest to want row {best}")
        # which output columns are right?
        cerr = (got - want).abs().amax(dim=(0, 1))[:64]
        print("   per-column max err (head 0):", [round(float(x), 3) for x in cerr.tolist()])


def bench_tma(B, H, N, iters=20):
    D = torch.device("cuda")
    C = H * 64
    qkv2 = ops.split_f16(torch.randn(B * N, 3 * C, device=D))
    for _ in range(3):
        ops.attention_packed_split(qkv2, B, N, H)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        ops.attention_packed_split(qkv2, B, N, H)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    print(f"bench attention TMA-fed B{B} H{H} L{N}: {ms:.3f} ms = {4.0 * B * H * N * N * 64 / ms / 1e9:.1f} algorithmic TFLOP/s", flush=True)


def trace_tma(B, H, N, n=512):
    """Per-phase clock stamps of the first `n` CTAs of one launch (scripts only: profiling aid)."""
    from edgecape_b200 import _lib
    D = torch.device("cuda")
    qkv2 = ops.split_f16(torch.randn(B * N, 3 * H * 64, device=D))
    for _ in range(3):
        ops.attention_packed_split(qkv2, B, N, H)
    buf = torch.zeros(n, 10, dtype=torch.int64, device=D)
    _lib.call("ec_attention_tc_set_trace", buf.data_ptr(), n)
    ops.attention_packed_split(qkv2, B, N, H)
    torch.cuda.synchronize()
    _lib.call("ec_attention_tc_set_trace", None, 0)
    t = buf.cpu().double()
    t = t[t[:, 9] > 0]                      # the persistent kernel runs one CTA per SM: only those rows are stamped
    d = t[:, 1:] - t[:, :-1]
    names = ["(persistent: start .. traced tile)", "Q/K TMA + S mma | persistent: wait for S", "row max", "P chunk 0",
             "P chunks (rest)", "drain PV", "epilogue", "final sync | persistent: remaining tiles", "dealloc"]
    tot = (t[:, 9] - t[:, 0])
    print(f"trace B{B} H{H} N{N}: CTA total mean {tot.mean():.0f} clk (min {tot.min():.0f} max {tot.max():.0f})")
    for i, nm in enumerate(names):
        print(f"  {nm:28s} mean {d[:, i].mean():8.0f}  min {d[:, i].min():8.0f}  max {d[:, i].max():8.0f}")
    first = t[:148, 0].min()
    print(f"  span of first {n} CTAs: {(t[:, 9].max() - first):.0f} clk")


if __name__ == "__main__":
    st = sys.argv[1]
    if "@" in st:                 # A/B: tma@1 = probabilities through shared memory, tma@2 = TMEM, narrow MMAs
        from edgecape_b200 import _lib
        st, v = st.split("@")
        _lib.call("ec_attention_tc_set_variant", int(v))
    if st == "tmabench":          # the ViT-B block attention alone (ncu target)
        bench_tma(32, 12, 325, iters=5)
        sys.exit(0)
    if st == "trace32":           # encoder self-attention of the head: 424 tokens, 8 heads x 32, key mask
        from edgecape_b200 import _lib
        D = torch.device("cuda")
        B, H, N, Dh = 16, 8, 424, 32
        qkv2 = ops.split_f16(torch.randn(B * N, 3 * H * Dh, device=D))
        mask = torch.zeros(B, N, dtype=torch.uint8, device=D)
        mask[:, 400:] = 1
        for masked in (True, False):
            km = mask if masked else None
            for _ in range(3):
                ops.attention_packed_split(qkv2, B, N, H, key_mask=km)
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            for _ in range(20):
                ops.attention_packed_split(qkv2, B, N, H, key_mask=km)
            e_.record()
            torch.cuda.synchronize()
            print(f"encoder-shaped attention masked={masked}: {s_.elapsed_time(e_) / 20 * 1e3:.1f} us")
            buf = torch.zeros(512, 10, dtype=torch.int64, device=D)
            _lib.call("ec_attention_tc_set_trace", buf.data_ptr(), 512)
            ops.attention_packed_split(qkv2, B, N, H, key_mask=km)
            torch.cuda.synchronize()
            _lib.call("ec_attention_tc_set_trace", None, 0)
            t = buf.cpu().double()
            d = t[:, 1:] - t[:, :-1]
            names = ["setup+tmem_alloc", "Q/K TMA + S mma", "row max", "P chunk 0", "P chunks (rest)", "drain PV",
                     "epilogue", "final sync", "dealloc"]
            print(f"  CTA total mean {(t[:, 9] - t[:, 0]).mean():.0f} clk")
            for i, nm in enumerate(names):
                print(f"  {nm:28s} mean {d[:, i].mean():8.0f}  min {d[:, i].min():8.0f}  max {d[:, i].max():8.0f}")
        sys.exit(0)
    if st == "trace":
        trace_tma(32, 12, 325)
        trace_tma(8, 16, 730)
        sys.exit(0)
    if st == "tma":
        check_tma(1, 64, 1)
        check_tma(1, 128, 1)
        check_tma(2, 325, 12)
        check_tma(1, 448, 2)
        check_tma(3, 17, 2)
        check_tma(2, 449, 3)
        check_tma(1, 730, 16)
        check_tma(1, 768, 2)
        bench_tma(8, 16, 730)
        bench_tma(32, 12, 325)
        bench(32, 12, 325)
        sys.exit(0)
    if st == "tiny":
        check(1, 1, 128, 64)
    elif st == "two":
        check(1, 2, 128, 128)
        check(1, 1, 100, 17)
    elif st == "vit":
        check(2, 12, 325, 325)
        check(1, 8, 100, 324)
        check(1, 8, 324, 100)
        check(1, 4, 130, 448)
    elif st == "bench":
        bench(32, 12, 325)
        bench(32, 16, 257)
