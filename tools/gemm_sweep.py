"""Development aid (not the bench contract): single-CTA vs CTA-pair (tcgen05 cta_group::2) GEMM kernels on the Qwen2-7B
projection shapes at M = 8192, bf16 / fp8 / W4A16, next to cuBLASLt (F.linear) and torch._scaled_mm.  TF/s from CUDA events."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xllm_b200 import ops  # noqa: E402

DEV, BF16 = "cuda", torch.bfloat16


def t_us(fn, it=6):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / it


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    shapes = {"qkv": (4608, 3584), "o": (3584, 3584), "gate_up": (37888, 3584), "down": (3584, 18944)}
    for name, (N, K) in shapes.items():
        a = torch.randn(M, K, device=DEV, dtype=BF16)
        w = torch.randn(N, K, device=DEV, dtype=BF16) * 0.02
        y = torch.empty(M, N, device=DEV, dtype=BF16)
        fl = 2.0 * M * N * K
        row = [f"{name:8s} {M}x{N}x{K}"]
        for mode, tag in ((1, "1cta"), (2, "pair")):
            ops.set_gemm_cta_pair(mode)
            row.append(f"bf16 {tag} {fl / t_us(lambda: ops.gemm_bf16(a, w, None, y)) / 1e6:7.1f}")
        y_ref = F.linear(a, w)
        ops.set_gemm_cta_pair(2)
        ops.gemm_bf16(a, w, None, y)
        row.append(f"maxdiff {float((y.float() - y_ref.float()).abs().max()):.3f}")
        row.append(f"cublas {fl / t_us(lambda: F.linear(a, w)) / 1e6:7.1f}")
        a8, w8 = a.to(torch.float8_e4m3fn), w.clamp(-1, 1).to(torch.float8_e4m3fn)
        one = torch.ones(1, device=DEV)
        for mode, tag in ((1, "1cta"), (2, "pair")):
            ops.set_gemm_cta_pair(mode)
            row.append(f"fp8 {tag} {fl / t_us(lambda: ops.cutlass_scaled_mm(y, a8, w8.t(), one, one, None)) / 1e6:7.1f}")
        row.append(f"scaled_mm {fl / t_us(lambda: torch._scaled_mm(a8, w8.t(), scale_a=one, scale_b=one, out_dtype=BF16)) / 1e6:7.1f}")
        del w, a8, w8
        qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (N // 16, K // 64, 32, 4), dtype=torch.int32, device=DEV)
        sc = (torch.rand(K // 128, N, device=DEV) * 0.01 + 0.001).to(BF16).view(torch.int16).to(torch.int32) & 0xFFFF
        meta = (sc | (0x4308 << 16)).contiguous()
        ys = []
        for mode, tag in ((1, "1cta"), (2, "pair")):
            ops.set_gemm_cta_pair(mode)
            row.append(f"w4 {tag} {fl / t_us(lambda: ops.gemm_w4a16(a, qw, meta, 128, None, y)) / 1e6:7.1f}")
            ys.append(y.clone())
        row.append("w4 pair==1cta " + str(bool(torch.equal(ys[0], ys[1]))))
        print(" | ".join(row), flush=True)
        del qw, meta, y, a
    ops.set_gemm_cta_pair(0)


if __name__ == "__main__":
    main()
