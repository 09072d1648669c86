"""Development probe: swap-AB FP8 GEMM at decode batch sizes with and without split-K (xb_set_gemm_splitk_workspace) on the
Llama-3-70B projections: unsharded, TP4 and TP8 shards.  us per call from CUDA-graph replay over rotating weight copies (> L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xllm_b200 import ops  # noqa: E402
from tools.fp8_swap_probe import graph_time, DEV, BF16, E4M3, PEAK  # noqa: E402


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    caps = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8]
    shapes = {}
    for tp in (8, 4, 1):
        shapes[f"qkv tp{tp}"] = (10240 // tp, 8192)
        shapes[f"o tp{tp}"] = (8192, 8192 // tp)
        shapes[f"gate_up tp{tp}"] = (57344 // tp, 8192)
        shapes[f"down tp{tp}"] = (8192, 28672 // tp)
    one = torch.ones(1, device=DEV)
    for name, (N, K) in shapes.items():
        copies = max(2, int(300e6 / (N * K)) + 1)
        ws = [torch.randn(N, K, device=DEV).clamp(-3, 3).to(E4M3) for _ in range(copies)]
        x = torch.randn(M, K, device=DEV).clamp(-3, 3).to(E4M3)
        y = torch.empty(M, N, device=DEV, dtype=BF16)
        nb = N * K + M * K + M * N * 2
        fns = [lambda i=i: ops.gemm_fp8_scaled(y, x, ws[i], one, one, None) for i in range(copies)]
        ops.disable_fp8_splitk()
        t0 = graph_time(fns)
        ops.gemm_fp8_scaled(y, x, ws[0], one, one, None)
        y0 = y.clone()
        line = f"{name:12s} N={N:6d} K={K:6d} M={M}: unsplit {t0:7.2f} us ({nb / t0 / 1e3 / PEAK:5.1%})"
        ops.enable_fp8_splitk(DEV)
        for cap in caps:
            ops.set_fp8_splitk_max(cap)
            t1 = graph_time(fns)
            ops.gemm_fp8_scaled(y, x, ws[0], one, one, None)
            err = float((y.float() - y0.float()).abs().max() / y0.float().abs().max())
            line += f"  split<={cap} {t1:7.2f} us ({nb / t1 / 1e3 / PEAK:5.1%}, diff {err:.1e})"
        ops.set_fp8_splitk_max(8)
        print(line, flush=True)
        del ws


if __name__ == "__main__":
    main()
