"""Checkpoint-side packing of weight-only quantised linears into the kernel layout.

Logical form (AWQ/GPTQ-like, oracle/quant.py spec): q uint8 [N,K] in 0..15, scales bf16 [N,K/g], zeros uint8 [N,K/g].
Kernel form: qweight int32 [N/16, K/64, 32, 4] (tile-packed nibbles, see csrc/linear_small_m.cu) and
meta int32 [K/g, N] = bf16(scale) | bf16(128+zero) << 16.  Packing runs on the host (C routine
xb_w4_pack_rows through ctypes, or the vectorised torch path below) - it is load-time work, not the hot path.
"""
import ctypes

import torch

from ._lib import check, lib


def pack_w4(q: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor, group_size: int = 128):
    """-> (qweight int32 [N/16,K/64,32,4], meta int32 [K/g,N]) on q's device."""
    N, K = q.shape
    assert N % 16 == 0 and K % 64 == 0 and K % group_size == 0 and group_size % 64 == 0
    qq = q.to(torch.int32).view(N // 16, 2, 8, K // 64, 4, 4, 4)          # [nt, half(r0/r1), g, kt, t, j, s]
    r0, r1 = qq[:, 0], qq[:, 1]                                           # [nt, g, kt, t, j, s]
    w = (r0[..., 0] | (r0[..., 1] << 16) | (r1[..., 0] << 4) | (r1[..., 1] << 20) |
         (r0[..., 2] << 8) | (r0[..., 3] << 24) | (r1[..., 2] << 12) | (r1[..., 3] << 28))   # [nt, g, kt, t, j]
    # int32 overflow of bit 31 wraps as intended
    qweight = w.permute(0, 2, 1, 3, 4).reshape(N // 16, K // 64, 32, 4).contiguous()
    s_bits = scales.to(torch.bfloat16).view(torch.int16).to(torch.int32) & 0xFFFF
    z_bits = (zeros.to(torch.float32) + 128.0).to(torch.bfloat16).view(torch.int16).to(torch.int32) & 0xFFFF
    meta = (s_bits | (z_bits << 16)).t().contiguous()
    return qweight, meta


def pack_w4_c(q: torch.Tensor) -> torch.Tensor:
    """Same nibble packing through the C routine (used by tests to pin the two packers against each other)."""
    N, K = q.shape
    qc = q.cpu().contiguous()
    out = torch.empty(N // 16, K // 64, 32, 4, dtype=torch.int32)
    check(lib().xb_w4_pack_rows(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(qc.data_ptr()), ctypes.c_int(N),
                                ctypes.c_int(K)), "w4_pack_rows")
    return out
