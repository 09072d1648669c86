"""Per-kernel device timings (CUDA events, rotating buffers larger than L2 so every launch streams from HBM).
Not the bench contract - a development aid; prints one line per kernel with achieved GB/s vs MEASURED_PEAKS."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xllm_b200 import ops, quant  # noqa: E402

DEV = "cuda"
BF16 = torch.bfloat16
PEAK = 6482.4
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, n_rot, iters=50, warm=5):
    for i in range(warm):
        fn(i % n_rot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_rot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def timeit_graph(fn, n_rot, iters=20):
    """all n_rot variants captured back to back in one graph: amortises launch overhead like the decode step."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(n_rot):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n_rot):
                fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / n_rot * 1e3


def report(name, us, nbytes, us_g=None):
    gbs = nbytes / us / 1e3
    extra = ""
    if us_g is not None:
        extra = f"  graph: {us_g:8.2f} us {nbytes / us_g / 1e3:8.1f} GB/s ({nbytes / us_g / 1e3 / PEAK:5.1%})"
    print(f"{name:44s} {us:8.2f} us {gbs:8.1f} GB/s ({gbs / PEAK:5.1%} of measured){extra}", flush=True)


def bench_decode(B, ctx, HQ=28, HKV=4, D=128, page=128, layers=28):
    npg = (ctx + page - 1) // page
    nblocks = B * npg + 1
    if os.environ.get("XB_MB_LAYOUT", "NHD") == "HND":     # head-major pages: each kv head's rows are contiguous
        mk = lambda: torch.randn(nblocks, HKV, page, D, device=DEV, dtype=BF16).permute(0, 2, 1, 3)
    else:                                                  # the reference layout [blocks, page, Hkv, D]
        mk = lambda: torch.randn(nblocks, page, HKV, D, device=DEV, dtype=BF16)
    caches = [(mk(), mk()) for _ in range(layers)]
    q = torch.randn(B, HQ, D, device=DEV, dtype=BF16)
    out = torch.empty_like(q)
    indptr = torch.arange(0, (B + 1) * npg, npg, dtype=torch.int32, device=DEV)
    indices = (torch.randperm(nblocks - 1, device=DEV) + 1).to(torch.int32)
    last = torch.full((B,), (ctx - 1) % page + 1, dtype=torch.int32, device=DEV)
    plan = ops.DecodePlan(B, HQ, HKV, D, page, npg, DEV)
    sc = 1 / math.sqrt(D)

    def fn(i):
        ops.batch_decode(plan, q, caches[i][0], caches[i][1], indptr, indices, last, sc, out)
    nbytes = 2 * B * ctx * HKV * D * 2 + 2 * B * HQ * D * 2 + 4 * B * npg
    us = timeit(fn, layers)
    us_g = timeit_graph(fn, layers)
    report(f"paged_decode B={B} ctx={ctx} chunk={plan.chunk_tokens} splits={plan.max_splits}", us, nbytes, us_g)


def bench_w4(N, K, M=1, gs=128, copies=None):
    copies = copies or max(2, int(300e6 / (N * K / 2)) + 1)
    ws = []
    for _ in range(copies):
        qw = torch.randint(-2**31, 2**31 - 1, (N // 16, K // 64, 32, 4), dtype=torch.int32, device=DEV)
        s = (torch.rand(K // gs, N, device=DEV) * 0.01 + 0.001).to(BF16).view(torch.int16).to(torch.int32) & 0xFFFF
        meta = s | (0x4308 << 16)
        ws.append((qw, meta.contiguous()))
    x = torch.randn(M, K, device=DEV, dtype=BF16)
    y = torch.empty(M, N, device=DEV, dtype=BF16)

    def fn(i):
        ops.w4a16_linear_small_m(x, ws[i][0], ws[i][1], gs, None, y)
    nbytes = N * K // 2 + (K // gs) * N * 4 + M * K * 2 + M * N * 2
    us = timeit(fn, copies)
    us_g = timeit_graph(fn, copies)
    report(f"w4a16 M={M} N={N} K={K}", us, nbytes, us_g)


def bench_bf16(N, K, M=1):
    copies = max(2, int(300e6 / (N * K * 2)) + 1)
    ws = [torch.randn(N, K, device=DEV, dtype=BF16) for _ in range(copies)]
    x = torch.randn(M, K, device=DEV, dtype=BF16)
    y = torch.empty(M, N, device=DEV, dtype=BF16)

    def fn(i):
        ops.matmul_small_m(x, ws[i], None, y)
    nbytes = N * K * 2 + M * K * 2 + M * N * 2
    report(f"bf16 linear M={M} N={N} K={K}", timeit(fn, copies), nbytes, timeit_graph(fn, copies))


def bench_gemm(kind, M, N, K, gs=128):
    E4M3 = torch.float8_e4m3fn
    if kind == "bf16":
        a = torch.randn(M, K, device=DEV, dtype=BF16)
        w = torch.randn(N, K, device=DEV, dtype=BF16)
        fn = lambda i: ops.gemm_bf16(a, w, None, y)
    elif kind == "fp8":
        a = torch.randn(M, K, device=DEV).to(E4M3)
        w = torch.randn(N, K, device=DEV).to(E4M3)
        s1 = torch.ones(1, device=DEV)
        fn = lambda i: ops.cutlass_scaled_mm(y, a, w.t(), s1, s1)
    elif kind == "w4":
        a = torch.randn(M, K, device=DEV, dtype=BF16)
        qw = torch.randint(-2**31, 2**31 - 1, (N // 16, K // 64, 32, 4), dtype=torch.int32, device=DEV)
        sb = (torch.rand(K // gs, N, device=DEV) * 0.01 + 0.001).to(BF16).view(torch.int16).to(torch.int32) & 0xFFFF
        meta = (sb | (0x4308 << 16)).contiguous()
        fn = lambda i: ops.gemm_w4a16(a, qw, meta, gs, None, y)
    elif kind == "torch":
        a = torch.randn(M, K, device=DEV, dtype=BF16)
        w = torch.randn(N, K, device=DEV, dtype=BF16)
        fn = lambda i: torch.matmul(a, w.t(), out=y)
    y = torch.empty(M, N, device=DEV, dtype=BF16)
    us = timeit(fn, 1, iters=20, warm=3)
    tf = 2.0 * M * N * K / us / 1e6
    print(f"gemm {kind:5s} M={M:6d} N={N:6d} K={K:6d}  {us:9.1f} us  {tf:8.1f} TFLOP/s", flush=True)


def bench_prefill(B, S, HQ=28, HKV=4, D=128):
    T = B * S
    qkv = torch.randn(T, (HQ + 2 * HKV) * D, device=DEV, dtype=BF16)
    q = qkv[:, :HQ * D].view(T, HQ, D)
    k = qkv[:, HQ * D:(HQ + HKV) * D].view(T, HKV, D)
    v = qkv[:, (HQ + HKV) * D:].view(T, HKV, D)
    cu = torch.arange(0, (B + 1) * S, S, dtype=torch.int32, device=DEV)
    out = torch.empty(T, HQ, D, device=DEV, dtype=BF16)
    fn = lambda i: ops.batch_prefill(q, k, v, cu, cu, D ** -0.5, out, None, max_qo_len=S)
    us = timeit(fn, 1, iters=10, warm=2)
    flops = 4.0 * HQ * D * S * S / 2 * B
    print(f"prefill attention B={B} S={S}: {us:9.1f} us  {flops / us / 1e6:7.1f} TFLOP/s (causal flops)", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1:
    # profiling entry: `microbench.py w4 N K M` | `microbench.py decode B ctx`
    if sys.argv[1] == "w4":
        bench_w4(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    elif sys.argv[1] == "decode":
        bench_decode(int(sys.argv[2]), int(sys.argv[3]))
    elif sys.argv[1] == "bf16":
        bench_bf16(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    elif sys.argv[1] == "prefill":
        for B, S in [(1, 2048), (4, 2048), (1, 8192), (16, 2048)]:
            bench_prefill(B, S)
    elif sys.argv[1] == "gemm":
        for kind in sys.argv[2].split(","):
            for M, N, K in [(2048, 4608, 3584), (2048, 37888, 3584), (2048, 3584, 18944), (8192, 37888, 3584), (8192, 8192, 8192)]:
                bench_gemm(kind, M, N, K)
    sys.exit(0)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "peak", PEAK)
    bench_decode(1, 4096)
    bench_decode(8, 4096)
    bench_decode(64, 4096)
    for N, K in [(4608, 3584), (3584, 3584), (37888, 3584), (3584, 18944)]:
        bench_w4(N, K, 1)
    bench_w4(37888, 3584, 8)
    bench_w4(37888, 3584, 64)
    bench_bf16(152064, 3584, 1)
    bench_bf16(4608, 3584, 1)
