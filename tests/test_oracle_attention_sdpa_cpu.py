"""The attention oracle (oracle/ops.py: ragged prefill, paged chunked prefill / decode over the block table) against an independent
public implementation on CPU: torch.nn.functional.scaled_dot_product_attention in fp32 per sequence, with the KV gathered from the
pages by plain indexing, GQA heads repeated, and the causal mask aligned to the END of the KV (query row i of a chunk of n rows over
L keys sees keys 0 .. L - n + i: FlashInfer's convention, flashinfer_attention.cpp:33-89).  This checks the page gathering, the ragged
indexing, the mask alignment and the GQA mapping of the oracle; the rounding ladder itself (bf16 P) is what the tolerance allows for
(tests/util.py: one implementation of the ladder sits 1.66e-3 relative L2 from exact math)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ops as O
from tests.util import assert_close_attention

BF16 = torch.bfloat16


def _sdpa(q, k, v, scale, causal):
    """q [n, HQ, D], k/v [L, HKV, D] -> [n, HQ, D] fp32; causal mask aligned to the end of the KV"""
    n, HQ, D = q.shape
    L, HKV, _ = k.shape
    rep = HQ // HKV
    qf = q.float().transpose(0, 1)[None]                                   # [1, HQ, n, D]
    kf = k.float().repeat_interleave(rep, dim=1).transpose(0, 1)[None]
    vf = v.float().repeat_interleave(rep, dim=1).transpose(0, 1)[None]
    mask = None
    if causal:
        i = torch.arange(n)[:, None]
        j = torch.arange(L)[None, :]
        mask = j <= (L - n + i)
    out = F.scaled_dot_product_attention(qf, kf, vf, attn_mask=mask, scale=scale)
    return out[0].transpose(0, 1)


@pytest.mark.parametrize("lens,HQ,HKV,D", [([5, 9, 1], 4, 2, 16), ([100, 257], 28, 4, 128), ([64], 8, 8, 64)])
def test_ragged_prefill_oracle_matches_sdpa(lens, HQ, HKV, D):
    g = torch.Generator().manual_seed(2026)
    T = sum(lens)
    q = torch.randn(T, HQ, D, generator=g).to(BF16)
    k = torch.randn(T, HKV, D, generator=g).to(BF16)
    v = torch.randn(T, HKV, D, generator=g).to(BF16)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    sc = 1.0 / math.sqrt(D)
    out = O.ragged_prefill_attention(q, k, v, cu, cu, sc, causal=True)
    ref = torch.cat([_sdpa(q[a:b], k[a:b], v[a:b], sc, True) for a, b in zip(cu[:-1].tolist(), cu[1:].tolist())])
    absv = torch.cat([_sdpa(q[a:b], k[a:b], v[a:b].abs(), sc, True) for a, b in zip(cu[:-1].tolist(), cu[1:].tolist())])
    # against EXACT math the elementwise bound is the worst case of one bf16 rounding of every p_j: 2^-8 * sum p|v| (= 3.9e-3)
    assert_close_attention(out, ref, absv, rtol=2.0 ** -8, what="ragged prefill oracle vs SDPA")


@pytest.mark.parametrize("q_lens,kv_lens,page,HQ,HKV,D,causal", [([1, 1, 1], [700, 33, 1], 16, 28, 4, 128, False),     # decode
                                                                  ([64, 3], [364, 50], 16, 28, 4, 128, True),           # chunked prefill
                                                                  ([7], [7], 4, 4, 2, 16, True)])                       # first chunk, paged
def test_paged_attention_oracle_matches_sdpa(q_lens, kv_lens, page, HQ, HKV, D, causal):
    g = torch.Generator().manual_seed(7)
    npg = [(n + page - 1) // page for n in kv_lens]
    nblocks = sum(npg) + 3
    kc = torch.randn(nblocks, page, HKV, D, generator=g).to(BF16)
    vc = torch.randn(nblocks, page, HKV, D, generator=g).to(BF16)
    perm = (torch.randperm(nblocks - 1, generator=g) + 1)[:sum(npg)].to(torch.int32)       # scattered physical blocks, block 0 unused
    indptr = torch.tensor([0] + list(torch.tensor(npg).cumsum(0)), dtype=torch.int32)
    last = torch.tensor([(n - 1) % page + 1 for n in kv_lens], dtype=torch.int32)
    qo = torch.tensor([0] + list(torch.tensor(q_lens).cumsum(0)), dtype=torch.int32)
    q = torch.randn(sum(q_lens), HQ, D, generator=g).to(BF16)
    sc = 1.0 / math.sqrt(D)
    out = O.paged_attention(q, kc, vc, qo, indptr, perm, last, sc, causal=causal)
    refs, absv = [], []
    for b, (n, L) in enumerate(zip(q_lens, kv_lens)):
        pages = perm[indptr[b]:indptr[b + 1]].long()
        k = kc[pages].reshape(-1, HKV, D)[:L]                               # gather by plain indexing
        v = vc[pages].reshape(-1, HKV, D)[:L]
        qb = q[qo[b]:qo[b + 1]]
        refs.append(_sdpa(qb, k, v, sc, causal))
        absv.append(_sdpa(qb, k, v.abs(), sc, causal))
    assert_close_attention(out, torch.cat(refs), torch.cat(absv), rtol=2.0 ** -8, what="paged attention oracle vs SDPA")
