"""Development aid (not the bench contract): sweeps the split / cluster geometry of the paged decode kernel on one shape
and prints device timings (isolated launches and PDL-chained back-to-back launches over per-layer caches > L2), checks
every variant against the cluster-less one, and dumps a per-CTA stage trace (xb_debug_set_decode_trace)."""
import ctypes
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xllm_b200 import _lib, ops  # noqa: E402

DEV, BF16 = "cuda", torch.bfloat16


def chained(fns, reps=5):
    for f in fns:
        f()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(20e6))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        for f in fns:
            f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * len(fns))


def isolated(fns, reps=3):
    ev = []
    torch.cuda._sleep(int(20e6))
    for r in range(reps):
        for f in fns:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record()
            if r:
                ev.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) * 1e3 / len(ev)


def graph_time(fns, iters=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fns:
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for f in fns:
                f()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters / len(fns)


def sweep(B, ctx, variants, HQ=28, HKV=4, D=128, page=128, layers=28, trace=False):
    npg = (ctx + page - 1) // page
    nblocks = B * npg + 1
    layers = max(2, min(layers, int(40e9 / (2 * nblocks * page * HKV * D * 2))))
    caches = [(torch.randn(nblocks, page, HKV, D, device=DEV, dtype=BF16), torch.randn(nblocks, page, HKV, D, device=DEV, dtype=BF16))
              for _ in range(layers)]
    q = torch.randn(B, HQ, D, device=DEV, dtype=BF16)
    indptr = torch.arange(0, (B + 1) * npg, npg, dtype=torch.int32, device=DEV)
    indices = (torch.randperm(nblocks - 1, device=DEV) + 1).to(torch.int32)
    last = torch.full((B,), (ctx - 1) % page + 1, dtype=torch.int32, device=DEV)
    sc = 1 / math.sqrt(D)
    nbytes = 2 * B * ctx * HKV * D * 2 + 2 * B * HQ * D * 2 + 4 * B * npg
    ref = None
    for name, env in variants:
        for k in ("XB_DECODE_CLUSTER", "XB_DECODE_CHUNK", "XB_DECODE_WARPS", "XB_DECODE_PARTS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        plan = ops.DecodePlan(B, HQ, HKV, D, page, npg, DEV)
        out = torch.empty_like(q)
        fns = [lambda i=i: ops.batch_decode(plan, q, caches[i][0], caches[i][1], indptr, indices, last, sc, out) for i in range(layers)]
        fns[0]()
        torch.cuda.synchronize()
        o0 = out.clone()
        if ref is None:
            ref = o0
        err = (o0.float() - ref.float()).abs().max().item()
        t_iso, t_ch, t_g = isolated(fns), chained(fns), graph_time(fns)
        print(f"B={B} ctx={ctx} {name:34s} splits={plan.max_splits:3d} cluster={plan.cluster:2d} chunk={plan.chunk_tokens:4d} | "
              f"isolated {t_iso:6.2f} us  chained {t_ch:6.2f} us ({nbytes / t_ch / 1e3:6.0f} GB/s)  graph {t_g:6.2f} us "
              f"({nbytes / t_g / 1e3:6.0f} GB/s) | max|diff vs first| {err:.2e}", flush=True)
        if trace:
            ncta = plan.max_splits * HKV * B
            buf = torch.zeros(ncta * 8, dtype=torch.int64, device=DEV)
            _lib.lib().xb_debug_set_decode_trace(ctypes.c_void_p(buf.data_ptr()))
            fns[1]()
            torch.cuda.synchronize()
            _lib.lib().xb_debug_set_decode_trace(ctypes.c_void_p(0))
            t = buf.view(ncta, 8).cpu()
            t0 = t[:, 0][t[:, 0] > 0].min().item()
            rows = []
            for i in (0, 1, 7, 2, 3, 4, 5, 6):
                col = t[:, i][t[:, i] > 0]
                if col.numel():
                    rows.append(f"s{i}: n={col.numel():3d} min {col.min().item() - t0:6d} med {int(col.median().item()) - t0:6d} max {col.max().item() - t0:6d}")
            print("    trace ns since first CTA start | " + " | ".join(rows), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "b1"
    print("max active clusters by cluster size:", {c: _lib.lib().xb_debug_max_active_clusters(c) for c in (2, 3, 4, 6, 8, 9, 10, 12, 16)}, flush=True)
    if which == "b1":
        sweep(1, 4096, [("default (37 partials, team merge)", {}),
                        ("chunk 128 (32 partials)", {"XB_DECODE_CHUNK": "128"}),
                        ("4 warps", {"XB_DECODE_WARPS": "4"}),
                        ("4 warps, chunk 64 (64 partials)", {"XB_DECODE_WARPS": "4", "XB_DECODE_CHUNK": "64"}),
                        ("1 x 16 cluster, chunk 256", {"XB_DECODE_CHUNK": "256", "XB_DECODE_CLUSTER": "16"}),
                        ("3 x 12 cluster", {"XB_DECODE_CLUSTER": "16"}),
                        ("2 x 8 cluster, chunk 256", {"XB_DECODE_CHUNK": "256", "XB_DECODE_CLUSTER": "8"})], trace=True)
        sweep(1, 8192, [("default", {}), ("1 x 16 cluster", {"XB_DECODE_CLUSTER": "16", "XB_DECODE_PARTS": "1"})], HQ=8, HKV=1)
    elif which == "batch":
        for B in (4, 8, 16, 32, 64):
            sweep(B, 4096, [("default", {}), ("cluster", {"XB_DECODE_CLUSTER": "16"})])
        sweep(32, 8192, [("default", {}), ("cluster", {"XB_DECODE_CLUSTER": "16"})], HQ=8, HKV=1)
