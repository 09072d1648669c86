"""Weight-only quantisation spec (W4A16 / W8A16).  The reference has NO weight-only kernel
(QuantArgs parses bits/group_size, xllm/core/framework/quant_args.h:36-60, nothing consumes them),
so this file DEFINES the arithmetic the CUDA kernels must reproduce - "parity unpinned".  Two forms of the same linear,
differing only in whether each dequantised weight is rounded to bf16 on its own:

  weights="bf16"   (tcgen05 GEMMs - the MMA operand has to be a bf16 tile - W8A16, and W4A16 with more than 8 tokens)
    w[n,k] = bf16( float(q[n,k] - z[n,k/g]) * float(s[n,k/g]) )        (one rounding)
    y      = bf16( sum_k float(x[m,k]) * float(w[n,k])  (+ bias) )     (fp32 accumulate)
    i.e. exactly "dequantise to the model dtype, then the reference's bf16 F::linear"
    (xllm/core/kernels/cuda/matmul.cpp:20-24), what the unquantised reference path computes on the dequantised
    checkpoint.

  weights="exact"  (W4A16 decode kernel with one token tile, M <= 8: scale / zero applied once per group AFTER the
                    integer dot product - 30 % fewer instructions in an issue-bound streaming loop)
    y      = bf16( sum_g s[n,g] * ( sum_{k in g} x[m,k] * (q[n,k] - z[n,g]) )  (+ bias) )   (fp32 accumulate)
    (q - z) * s is exact in fp32 (5 x 8 significant bits), so this is the same sum with UN-rounded weights; the two forms
    differ by the 2^-8-relative rounding of each individual weight (rel-L2 ~1.7e-3 of the output on random data).

q, z are unsigned (4 bit: 0..15; 8 bit: 0..255), s is bf16, g = group_size along K (AWQ/GPTQ style; GPTQ sym =>
z = 2^(bits-1)).  `w4a16_form(M)` names the form the library computes for a given token count.
TEST INFRASTRUCTURE - see oracle/__init__.py.
"""
import torch

BF16 = torch.bfloat16
F32 = torch.float32


def quantize(w: torch.Tensor, bits: int = 4, group_size: int = 128, sym: bool = False):
    """bf16/fp32 W[N,K] -> (q uint8 [N,K], scales bf16 [N,K/g], zeros uint8 [N,K/g]).  min/max asymmetric
    (AWQ-like) or symmetric with z = 2^(bits-1) (GPTQ sym)."""
    N, K = w.shape
    qmax = (1 << bits) - 1
    wg = w.to(F32).view(N, K // group_size, group_size)
    if sym:
        amax = wg.abs().amax(-1, keepdim=True).clamp_min(1e-8)
        scale = (amax / (qmax // 2)).to(BF16).to(F32)
        zero = torch.full_like(scale, float((qmax + 1) // 2))
    else:
        wmin = wg.amin(-1, keepdim=True).clamp_max(0)
        wmax = wg.amax(-1, keepdim=True).clamp_min(0)
        scale = ((wmax - wmin).clamp_min(1e-8) / qmax).to(BF16).to(F32)
        zero = torch.round(-wmin / scale).clamp(0, qmax)
    q = torch.clamp(torch.round(wg / scale) + zero, 0, qmax)
    return (q.view(N, K).to(torch.uint8), scale.squeeze(-1).to(BF16), zero.squeeze(-1).to(torch.uint8))


def dequantize(q, scales, zeros, group_size: int = 128) -> torch.Tensor:
    """-> bf16 W[N,K] per the spec above."""
    N, K = q.shape
    qf = q.to(F32).view(N, K // group_size, group_size)
    w = (qf - zeros.to(F32).unsqueeze(-1)) * scales.to(F32).unsqueeze(-1)
    return w.view(N, K).to(BF16)


def dequantize_exact(q, scales, zeros, group_size: int = 128) -> torch.Tensor:
    """-> fp32 W[N,K] = (q - z) * s, exact (no rounding: 5-bit integer x 8-bit significand)."""
    N, K = q.shape
    qf = q.to(F32).view(N, K // group_size, group_size)
    w = (qf - zeros.to(F32).unsqueeze(-1)) * scales.to(F32).unsqueeze(-1)
    return w.view(N, K)


def w4a16_form(M: int) -> str:
    """the form of the W4A16 linear the library computes for M tokens (see the module docstring)."""
    return "exact" if M <= 8 else "bf16"


def linear_wna16(x, q, scales, zeros, group_size=128, bias=None, weights="bf16"):
    if weights not in ("bf16", "exact"):
        raise ValueError(weights)
    w = dequantize(q, scales, zeros, group_size) if weights == "bf16" else dequantize_exact(q, scales, zeros, group_size)
    y = x.to(F32) @ w.to(F32).t()
    if bias is not None:
        y = y + bias.to(F32)
    return y.to(BF16)
