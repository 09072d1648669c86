"""Regression pin of the oracle: tests/golden/oracle_fixtures.pt holds frozen seeded input -> output pairs of every oracle
op family (written by tests/golden/make_oracle_fixtures.py).  Elementwise / integer results must reproduce bit-for-bit;
results that pass through a CPU matmul (attention, linears) within one bf16 ulp (BLAS blocking may differ between hosts).
"""
import math
import os

import pytest
import torch

from oracle import ops as O
from oracle import quant as Q
from tests.util import assert_close_bf16

BF16 = torch.bfloat16
PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_fixtures.pt")


@pytest.fixture(scope="module")
def fx():
    return torch.load(PATH, weights_only=False)


def test_fixture_file_is_fresh(fx):
    """the generator, run now, produces the same case set (a new case must be regenerated AND committed)."""
    from tests.golden.make_oracle_fixtures import cases
    assert sorted(cases().keys()) == sorted(fx.keys())


def test_elementwise_ops_bit_exact(fx):
    x, w, eps = fx["rms_norm"]["inp"]
    assert torch.equal(O.rms_norm(x, w, eps), fx["rms_norm"]["out"])
    x, res, w, eps = fx["fused_add_rms_norm"]["inp"]
    h, r = O.fused_add_rms_norm(x, res, w, eps)
    assert torch.equal(h, fx["fused_add_rms_norm"]["out"][0]) and torch.equal(r, fx["fused_add_rms_norm"]["out"][1])
    pos, q, k, rot, maxpos, theta = fx["rotary_embedding"]["inp"]
    cs = O.compute_cos_sin_cache(rot, maxpos, theta, BF16)
    q2, k2 = O.rotary_embedding(pos, q, k, cs, is_neox=True)
    assert torch.equal(q2, fx["rotary_embedding"]["out"][0]) and torch.equal(k2, fx["rotary_embedding"]["out"][1])
    (gu,) = fx["act_and_mul_silu"]["inp"]
    assert torch.equal(O.act_and_mul(gu, "silu"), fx["act_and_mul_silu"]["out"])
    (xf,) = fx["fp8_scaled_quantize"]["inp"]
    x8, sc = O.fp8_scaled_quantize(xf)
    assert torch.equal(x8.view(torch.uint8), fx["fp8_scaled_quantize"]["out"][0])
    assert torch.equal(sc, fx["fp8_scaled_quantize"]["out"][1])


def test_quantiser_bit_exact_and_linear(fx):
    wq, xl, bl = fx["w4a16_linear"]["inp"]
    qq, ss, zz, y = fx["w4a16_linear"]["out"]
    q2, s2, z2 = Q.quantize(wq, 4, 128)
    assert torch.equal(q2, qq) and torch.equal(s2, ss) and torch.equal(z2, zz)
    assert_close_bf16(Q.linear_wna16(xl, qq, ss, zz, 128, bl), y, ulps=1, rel_l2=1e-4, what="w4a16 linear fixture")


def test_attention_fixtures(fx):
    qd, kc, vc, indptr, perm, last, page = fx["paged_decode"]["inp"]
    D = qd.shape[-1]
    qo = torch.arange(qd.shape[0] + 1, dtype=torch.int32)
    o, lse = O.paged_attention(qd, kc, vc, qo, indptr, perm, last, 1 / math.sqrt(D), causal=False, return_lse=True)
    assert_close_bf16(o, fx["paged_decode"]["out"][0], ulps=1, rel_l2=1e-4, what="paged decode fixture")
    assert torch.allclose(lse, fx["paged_decode"]["out"][1], rtol=1e-5, atol=1e-5)
    qp, kp, vp, cu = fx["ragged_prefill"]["inp"]
    o = O.ragged_prefill_attention(qp, kp, vp, cu, cu, 1 / math.sqrt(qp.shape[-1]))
    assert_close_bf16(o, fx["ragged_prefill"]["out"], ulps=1, rel_l2=1e-4, what="ragged prefill fixture")


@pytest.mark.gpu
def test_kernels_against_frozen_outputs(fx, built_lib):
    """the CUDA kernels against the FROZEN outputs (not against a fresh oracle evaluation)."""
    from xllm_b200 import ops
    dev = "cuda"
    x, w, eps = fx["rms_norm"]["inp"]
    out = torch.empty_like(x, device=dev)
    ops.rms_norm(out, x.to(dev), w.to(dev), eps)
    assert_close_bf16(out, fx["rms_norm"]["out"], ulps=1, what="rms_norm kernel vs fixture")
    (gu,) = fx["act_and_mul_silu"]["inp"]
    out = torch.empty(gu.shape[0], gu.shape[1] // 2, dtype=BF16, device=dev)
    ops.act_and_mul(out, gu.to(dev), "silu")
    assert_close_bf16(out, fx["act_and_mul_silu"]["out"], ulps=1, what="act_and_mul kernel vs fixture")
    qd, kc, vc, indptr, perm, last, page = fx["paged_decode"]["inp"]
    B, HQ, D = qd.shape
    plan = ops.DecodePlan(B, HQ, kc.shape[2], D, page, int((indptr[1:] - indptr[:-1]).max()), dev)
    o = torch.empty(B, HQ, D, dtype=BF16, device=dev)
    ops.batch_decode(plan, qd.to(dev), kc.to(dev), vc.to(dev), indptr.to(dev), perm.to(dev), last.to(dev), 1 / math.sqrt(D), o, None)
    # attention: forward-error bound of tests/util.py (bf16 P roundings; elements that cancel to ~0 cannot be held to ulps)
    from tests.util import assert_close_attention
    qo = torch.arange(B + 1, dtype=torch.int32)
    scale = O.paged_attention(qd, kc, vc.abs(), qo, indptr, perm, last, 1 / math.sqrt(D), causal=False)
    assert_close_attention(o, fx["paged_decode"]["out"][0], scale, what="paged decode kernel vs fixture")
