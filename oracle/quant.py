"""Weight-only quantisation spec (W4A16 / W8A16).  The reference has NO weight-only kernel
(QuantArgs parses bits/group_size, xllm/core/framework/quant_args.h:36-60, nothing consumes them),
so this file DEFINES the arithmetic the CUDA kernels must reproduce - "parity unpinned":

    w[n,k] = bf16( float(q[n,k] - z[n,k/g]) * float(s[n,k/g]) )        (one rounding)
    y      = bf16( sum_k float(x[m,k]) * float(w[n,k])  (+ bias) )     (fp32 accumulate)

Second form ("exact", what the decode-sized W4A16 kernels may compute - xb_set_w4_decode_form):

    y      = bf16( sum_g float(s[n,g]) * ( sum_{k in g} float(x[m,k]) * (q[n,k] - z[n,g]) )  (+ bias) )

the same dequantisation WITHOUT the intermediate bf16 rounding of every weight: form 1 perturbs each product by at
most 2^-9 relative (the rounding of w), so |y_exact - y_bf16w| <= 2^-9 * sum_k |x||w| and, for independent rounding
errors, ~2^-9 * sqrt(sum_k (x w)^2) - the size of one bf16 ulp of a typical output.

Form 1 is exactly "dequantise to the model dtype, then the reference's bf16 F::linear"
(xllm/core/kernels/cuda/matmul.cpp:20-24), which is what the unquantised reference path computes
on the dequantised checkpoint.  q, z are unsigned (4 bit: 0..15; 8 bit: 0..255), s is bf16,
g = group_size along K (AWQ/GPTQ style; GPTQ sym => z = 2^(bits-1)).
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


def linear_wna16(x, q, scales, zeros, group_size=128, bias=None, form="bf16w"):
    """form "bf16w": weights rounded to bf16 first (spec form 1); "exact": scale / zero applied to the integer dot
    product of every group (spec form 2, float64 here so the checker itself adds no rounding)."""
    if form == "exact":
        N, K = q.shape
        M = x.shape[0]
        xg = x.to(torch.float64).view(M, K // group_size, group_size)
        qg = (q.to(torch.float64).view(N, K // group_size, group_size) - zeros.to(torch.float64).unsqueeze(-1))
        y = torch.einsum("mgk,ngk,ng->mn", xg, qg, scales.to(torch.float64))
        if bias is not None:
            y = y + bias.to(torch.float64)
        return y.to(F32).to(BF16)
    assert form == "bf16w", form
    w = dequantize(q, scales, zeros, group_size)
    y = x.to(F32) @ w.to(F32).t()
    if bias is not None:
        y = y + bias.to(F32)
    return y.to(BF16)
