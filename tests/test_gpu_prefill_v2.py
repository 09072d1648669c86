"""GPU parity of the ping-pong prefill attention kernel (two q tiles per CTA, two softmax warpgroups;
xb_set_prefill_variant(1)): the ragged / paged / decode-on-tensor-cores cases of tests/test_gpu_prefill.py re-run under
that variant against the same oracle, plus bit-identity with the one-tile-per-CTA kernel (same arithmetic ladder, same
kv tile order per row => identical results)."""
import math

import pytest
import torch

from tests import test_gpu_prefill as P

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
DEV = "cuda"


@pytest.fixture
def v2(built_lib):
    from xllm_b200 import ops
    old = ops.set_prefill_variant(1)
    yield
    ops.set_prefill_variant(old)


@pytest.mark.parametrize("lens,HQ,HKV,D", P.RAGGED + [([2048, 2048], 28, 4, 128), ([4096], 32, 8, 128)])
def test_batch_prefill_ragged_v2(lens, HQ, HKV, D, v2, built_lib):
    P.test_batch_prefill_ragged(lens, HQ, HKV, D, built_lib)


@pytest.mark.parametrize("q_lens,kv_lens,HQ,HKV,D,page,causal", P.PAGED)
def test_batch_chunked_prefill_paged_v2(q_lens, kv_lens, HQ, HKV, D, page, causal, v2, built_lib):
    P.test_batch_chunked_prefill_paged(q_lens, kv_lens, HQ, HKV, D, page, causal, built_lib)


def test_prefill_decode_kernels_agree_v2(v2, built_lib):
    P.test_prefill_decode_kernels_agree(built_lib)


@pytest.mark.parametrize("lens,HQ,HKV,D", [([2048], 28, 4, 128), ([1, 5, 300, 17], 28, 4, 128), ([333, 700], 14, 2, 64)])
def test_variants_bit_identical(lens, HQ, HKV, D, built_lib):
    from xllm_b200 import ops
    qkv, cu = P._ragged_case(lens, HQ, HKV, D)
    T = sum(lens)
    qd = qkv.to(DEV)
    q = qd[:, :HQ * D].view(T, HQ, D)
    k = qd[:, HQ * D:(HQ + HKV) * D].view(T, HKV, D)
    v = qd[:, (HQ + HKV) * D:].view(T, HKV, D)
    cud = cu.to(DEV)
    outs, lses = [], []
    for variant in (0, 1):
        old = ops.set_prefill_variant(variant)
        o = torch.empty(T, HQ, D, dtype=BF16, device=DEV)
        lse = torch.empty(T, HQ, dtype=torch.float32, device=DEV)
        ops.batch_prefill(q, k, v, cud, cud, 1.0 / math.sqrt(D), o, lse, max_qo_len=max(lens))
        torch.cuda.synchronize()
        ops.set_prefill_variant(old)
        outs.append(o)
        lses.append(lse)
    assert torch.equal(outs[0], outs[1]), "ping-pong kernel output differs from the one-tile kernel"
    assert torch.equal(lses[0], lses[1]), "LSE differs"
