"""GPU parity: paged decode attention through the C ABI vs the oracle."""
import math

import pytest
import torch

from oracle import ops as O
from tests.util import assert_close_attention, assert_close_bf16

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
DEV = "cuda"


def make_case(kv_lens, HQ, HKV, D, page, seed=2026, extra_blocks=7):
    """KV cache N(0,1) bf16, pages = random permutation of physical blocks, block 0 reserved."""
    g = torch.Generator().manual_seed(seed)
    B = len(kv_lens)
    npages = [(n + page - 1) // page for n in kv_lens]
    total = sum(npages)
    nblocks = total + 1 + extra_blocks
    perm = (torch.randperm(nblocks - 1, generator=g) + 1)[:total].to(torch.int32)
    indptr = torch.tensor([0] + list(torch.tensor(npages).cumsum(0).tolist()), dtype=torch.int32)
    last = torch.tensor([(n - 1) % page + 1 if n > 0 else 0 for n in kv_lens], dtype=torch.int32)
    kc = torch.randn(nblocks, page, HKV, D, generator=g).to(BF16)
    vc = torch.randn(nblocks, page, HKV, D, generator=g).to(BF16)
    q = torch.randn(B, HQ, D, generator=g).to(BF16)
    return q, kc, vc, indptr, perm, last


def run_gpu(q, kc, vc, indptr, indices, last, page, max_pages, lse=False, num_sms=None):
    from xllm_b200 import ops
    B, HQ, D = q.shape
    plan = ops.DecodePlan(B, HQ, kc.shape[2], D, page, max_pages, DEV, num_sms=num_sms)
    out = torch.empty(B, HQ, D, dtype=BF16, device=DEV)
    lse_t = torch.empty(B, HQ, dtype=torch.float32, device=DEV) if lse else None
    kcd, vcd = kc.to(DEV), vc.to(DEV)
    args = (q.to(DEV), kcd, vcd, indptr.to(DEV), indices.to(DEV), last.to(DEV), 1.0 / math.sqrt(D), out, lse_t)
    ops.batch_decode(plan, *args)
    first = out.clone()
    ops.batch_decode(plan, *args)          # second launch: the ticket counters must have been restored
    torch.cuda.synchronize()
    assert torch.equal(first, out), "relaunch with the same workspace changed the result"
    return out, lse_t, plan


CASES = [
    # kv_lens,                 HQ, HKV, D,  page
    ([4096],                   28, 4, 128, 128),   # BASELINE configs[1] attention shape
    ([1],                      28, 4, 128, 128),
    ([17, 700, 1, 129, 2048],  28, 4, 128, 16),    # ragged batch, small pages
    ([333, 64],                14, 2, 64, 16),     # Qwen2-0.5B heads
    ([257] * 4,                16, 8, 128, 16),    # the MLU decode KAT shape
    ([1000, 31],               8, 8, 128, 32),     # MHA (group 1)
    ([777],                    32, 2, 128, 64),    # group 16 (two row halves)
    ([500, 3],                 64, 2, 64, 128),    # group 32 -> two head tiles
    ([45, 46, 47],             8, 1, 128, 1),      # page_size 1
    ([100, 260],               12, 4, 64, 12),     # non power-of-two page size
    ([8192],                   8, 1, 128, 128),    # Llama-3-70B TP8 per-GPU heads, ctx 8192
    ([133],                    14, 2, 64, 128),    # BASELINE configs[0] heads a few tokens into the second page
    ([4096, 9, 300],           28, 4, 128, 128),   # ragged lengths under one split count
]


@pytest.mark.parametrize("kv_lens,HQ,HKV,D,page", CASES)
def test_paged_decode_matches_oracle(kv_lens, HQ, HKV, D, page, built_lib):
    q, kc, vc, indptr, indices, last = make_case(kv_lens, HQ, HKV, D, page)
    B = len(kv_lens)
    qo = torch.arange(B + 1, dtype=torch.int32)
    ref, ref_lse = O.paged_attention(q, kc, vc, qo, indptr, indices, last, 1.0 / math.sqrt(D), causal=False,
                                     return_lse=True)
    max_pages = max((n + page - 1) // page for n in kv_lens)
    out, lse, plan = run_gpu(q, kc, vc, indptr, indices, last, page, max_pages, lse=True)
    scale = O.paged_attention(q, kc, vc.abs(), qo, indptr, indices, last, 1.0 / math.sqrt(D), causal=False)
    assert_close_attention(out, ref, scale, what=f"paged_decode {kv_lens} splits={plan.max_splits}")
    # l sums bf16-rounded P relative to the running maximum of the tile order: log2(l) moves by ~2^-9 / ln 2
    assert torch.allclose(lse.cpu(), ref_lse, rtol=1e-4, atol=5e-3), "base-2 LSE mismatch"


def test_paged_decode_distance_from_exact_math(built_lib):
    """how far ONE implementation of the bf16-P ladder sits from un-rounded softmax numerators (oracle with fp32 P,
    same bf16 output rounding): analytically 1.66e-3 relative L2 from the rounding of P alone (tests/util.py), plus the
    partly independent bf16 roundings of the two outputs.  Asserted for the kernel AND for the reference-ladder oracle,
    at the BASELINE configs[1] attention shape."""
    kv_lens, HQ, HKV, D, page = [4096], 28, 4, 128, 128
    q, kc, vc, indptr, indices, last = make_case(kv_lens, HQ, HKV, D, page)
    qo = torch.arange(2, dtype=torch.int32)
    sc = 1.0 / math.sqrt(D)
    exact = O.paged_attention(q, kc, vc, qo, indptr, indices, last, sc, causal=False, round_p=False)
    ladder = O.paged_attention(q, kc, vc, qo, indptr, indices, last, sc, causal=False)
    out, _, _ = run_gpu(q, kc, vc, indptr, indices, last, page, 32)
    scale = O.paged_attention(q, kc, vc.abs(), qo, indptr, indices, last, sc, causal=False)
    assert_close_attention(out, exact, scale, what="paged_decode kernel vs fp32-P oracle (one bf16-P implementation)", rel_l2=2.6e-3)
    assert_close_attention(ladder, exact, scale, what="reference-ladder oracle vs fp32-P oracle", rel_l2=2.6e-3)


def test_paged_decode_upper_bound_plan(built_lib):
    """CUDA-graph style: the plan is made for a larger page budget than any request uses; CTAs past the live
    split count must exit and the result must not change."""
    kv_lens = [300, 1200]
    q, kc, vc, indptr, indices, last = make_case(kv_lens, 28, 4, 128, 16)
    qo = torch.arange(3, dtype=torch.int32)
    ref = O.paged_attention(q, kc, vc, qo, indptr, indices, last, 1.0 / math.sqrt(128), causal=False)
    out, _, plan = run_gpu(q, kc, vc, indptr, indices, last, 16, max_pages=4096 // 16)
    assert plan.max_splits > 1
    scale = O.paged_attention(q, kc, vc.abs(), qo, indptr, indices, last, 1.0 / math.sqrt(128), causal=False)
    assert_close_attention(out, ref, scale, what="paged_decode upper-bound plan")


def test_paged_decode_empty_and_padding_rows(built_lib):
    """padded decode rows of the reference use block 0 / last_page_len 1 (batch_input_builder.cpp:854-873)."""
    g = torch.Generator().manual_seed(1)
    HQ, HKV, D, page = 28, 4, 128, 128
    kc = torch.randn(8, page, HKV, D, generator=g).to(BF16)
    vc = torch.randn(8, page, HKV, D, generator=g).to(BF16)
    q = torch.randn(3, HQ, D, generator=g).to(BF16)
    indptr = torch.tensor([0, 2, 3, 4], dtype=torch.int32)
    indices = torch.tensor([5, 2, 0, 0], dtype=torch.int32)
    last = torch.tensor([40, 1, 1], dtype=torch.int32)
    qo = torch.arange(4, dtype=torch.int32)
    ref = O.paged_attention(q, kc, vc, qo, indptr, indices, last, 1.0 / math.sqrt(D), causal=False)
    out, _, _ = run_gpu(q, kc, vc, indptr, indices, last, page, 2)
    scale = O.paged_attention(q, kc, vc.abs(), qo, indptr, indices, last, 1.0 / math.sqrt(D), causal=False)
    assert_close_attention(out, ref, scale, what="padding rows")
    # a row that attends to exactly one token returns that token's V bit-exactly
    assert torch.equal(out[1].cpu(), vc[0, 0].repeat_interleave(HQ // HKV, 0))


def test_paged_decode_properties_full_size(built_lib):
    """BASELINE config sizes where the per-element oracle is slow: size-independent properties.
    (a) V = const c  =>  output == c exactly for every head; (b) permuting physical pages (and the table with
    them) does not change the result bit-for-bit; (c) appending a -inf-score-equivalent is not available, so
    instead: splitting the same problem with a different SM budget (different split count) stays within 1 ulp."""
    B, HQ, HKV, D, page, ctx = 64, 28, 4, 128, 128, 4096
    g = torch.Generator().manual_seed(7)
    npg = ctx // page
    nblocks = B * npg + 1
    kc = torch.randn(nblocks, page, HKV, D, generator=g).to(BF16).to(DEV)
    vc = torch.full((nblocks, page, HKV, D), 0.375, dtype=BF16, device=DEV)
    q = torch.randn(B, HQ, D, generator=g).to(BF16)
    indptr = torch.arange(0, (B + 1) * npg, npg, dtype=torch.int32)
    perm = (torch.randperm(nblocks - 1, generator=g) + 1).to(torch.int32)
    last = torch.full((B,), page, dtype=torch.int32)
    out, _, _ = run_gpu(q, kc, vc, indptr, perm, last, page, npg)
    assert torch.all(out == 0.375), "convex combination of a constant V must return the constant"
    # (b) page permutation invariance
    vc2 = torch.randn(nblocks, page, HKV, D, generator=g).to(BF16).to(DEV)
    out1, _, _ = run_gpu(q, kc, vc2, indptr, perm, last, page, npg)
    shuffle = torch.randperm(nblocks - 1, generator=g) + 1
    inv = torch.zeros(nblocks, dtype=torch.long)
    inv[shuffle] = torch.arange(1, nblocks)
    kc3, vc3 = kc.clone(), vc2.clone()
    kc3[1:] = kc[shuffle.to(DEV)]
    vc3[1:] = vc2[shuffle.to(DEV)]
    perm3 = inv[perm.long()].to(torch.int32)
    out2, _, _ = run_gpu(q, kc3, vc3, indptr, perm3, last, page, npg)
    assert torch.equal(out1, out2), "result depends on physical page placement"
    # (c) different split factor
    out3, _, _ = run_gpu(q, kc, vc2, indptr, perm, last, page, npg, num_sms=1024)
    qo = torch.arange(B + 1, dtype=torch.int32)
    assert_close_attention(out3, out1, torch.full(out1.shape, 0.8 * 0.5), what="split-count independence")
