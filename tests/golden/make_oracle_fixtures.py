"""Freezes small seeded input -> output pairs of the ORACLE (oracle/*.py) as tests/golden/oracle_fixtures.pt.

Purpose: a regression pin of the checker itself.  The GPU parity tests compare the CUDA kernels with the oracle on the
same seeded inputs; this file guarantees that an edit of the oracle cannot silently move the target - the CPU suite
(tests/test_oracle_fixtures.py) re-evaluates every case bit-for-bit, and the opt-in GPU part of that test runs the
kernels against the frozen outputs directly.  These are NOT reference outputs (the reference cannot run here, DESIGN.md
section 2); the reference-held golden vectors are the ones in tests/test_oracle_golden.py.

    python tests/golden/make_oracle_fixtures.py        # rewrites oracle_fixtures.pt (commit the result)
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ops as O      # noqa: E402
from oracle import quant as Q    # noqa: E402

BF16 = torch.bfloat16


def cases():
    g = torch.Generator().manual_seed(2026)
    rn = lambda *s, std=1.0: (torch.randn(*s, generator=g) * std).to(BF16)
    out = {}
    # RMSNorm / fused add (norm.cu rounding order)
    x, res, w = rn(5, 256), rn(5, 256), (1 + 0.1 * torch.randn(256, generator=g)).to(BF16)
    out["rms_norm"] = dict(inp=(x, w, 1e-6), out=O.rms_norm(x, w, 1e-6))
    h, r = O.fused_add_rms_norm(x, res, w, 1e-6)
    out["fused_add_rms_norm"] = dict(inp=(x, res, w, 1e-6), out=(h, r))
    # RoPE (neox) with the reference's cos/sin cache
    cs = O.compute_cos_sin_cache(64, 128, 1e6, BF16)
    pos = torch.tensor([0, 1, 17, 127, 64])
    q, k = rn(5, 4, 64), rn(5, 2, 64)
    q2, k2 = O.rotary_embedding(pos, q, k, cs, is_neox=True)
    out["rotary_embedding"] = dict(inp=(pos, q, k, 64, 128, 1e6), out=(q2, k2))
    # SiLU * mul
    gu = rn(3, 512)
    out["act_and_mul_silu"] = dict(inp=(gu,), out=O.act_and_mul(gu, "silu"))
    # paged decode attention: ragged batch, GQA 7, scattered pages
    page, HQ, HKV, D = 16, 14, 2, 64
    kv_lens = [1, 37, 64]
    npg = [(n + page - 1) // page for n in kv_lens]
    nblocks = sum(npg) + 3
    perm = (torch.randperm(nblocks - 1, generator=g) + 1)[: sum(npg)].to(torch.int32)
    indptr = torch.tensor([0] + torch.tensor(npg).cumsum(0).tolist(), dtype=torch.int32)
    last = torch.tensor([(n - 1) % page + 1 for n in kv_lens], dtype=torch.int32)
    kc, vc, qd = rn(nblocks, page, HKV, D), rn(nblocks, page, HKV, D), rn(3, HQ, D)
    qo = torch.arange(4, dtype=torch.int32)
    o, lse = O.paged_attention(qd, kc, vc, qo, indptr, perm, last, 1 / math.sqrt(D), causal=False, return_lse=True)
    out["paged_decode"] = dict(inp=(qd, kc, vc, indptr, perm, last, page), out=(o, lse))
    # ragged causal prefill
    lens = [5, 33]
    cu = torch.tensor([0, 5, 38], dtype=torch.int32)
    qp, kp, vp = rn(38, HQ, D), rn(38, HKV, D), rn(38, HKV, D)
    out["ragged_prefill"] = dict(inp=(qp, kp, vp, cu), out=O.ragged_prefill_attention(qp, kp, vp, cu, cu, 1 / math.sqrt(D)))
    # W4A16 linear (spec of oracle/quant.py) incl. the quantiser
    wq = rn(48, 256, std=0.02)
    qq, ss, zz = Q.quantize(wq, 4, 128)
    xl, bl = rn(3, 256), rn(48)
    out["w4a16_linear"] = dict(inp=(wq, xl, bl), out=(qq, ss, zz, Q.linear_wna16(xl, qq, ss, zz, 128, bl)))
    # fp8 dynamic / static quant + scaled matmul
    xf = rn(4, 128, std=3.0)
    x8, sc = O.fp8_scaled_quantize(xf)
    out["fp8_scaled_quantize"] = dict(inp=(xf,), out=(x8.view(torch.uint8), sc))
    return out


if __name__ == "__main__":
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_fixtures.pt")
    torch.save(cases(), dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")
