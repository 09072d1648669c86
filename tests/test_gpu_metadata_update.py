"""GPU parity (bit-exact, integer work): the CUDA-graph decode metadata refresh (xb_decode_metadata_update) against the CPU
restatement of llm_decode_metadata_update_kernel (xllm/core/kernels/cuda/llm_decode_metadata_update.cu:29-62), including the
padding rule and the untouched tails of the persistent buffers; and a graph replay that serves a changed batch without a
host-side plan."""
import math

import pytest
import torch

from oracle import batch as OB

pytestmark = pytest.mark.gpu
DEV = "cuda"
I32 = torch.int32


def _case(n_tok, padded, batch, n_idx, cap_tok, cap_batch, cap_idx, seed):
    g = torch.Generator().manual_seed(seed)
    ri = lambda n, hi=100000: torch.randint(0, hi, (n,), generator=g, dtype=I32)
    kv = torch.cat([torch.zeros(1, dtype=I32), ri(batch, 500).cumsum(0).to(I32)])
    src = dict(tokens=ri(n_tok), positions=ri(n_tok), new_cache_slots=ri(n_tok), kv_seq_lens=kv,
               paged_kv_indptr=torch.cat([torch.zeros(1, dtype=I32), ri(batch, 40).cumsum(0).to(I32)]),
               paged_kv_indices=ri(max(n_idx, 1)), paged_kv_last_page_len=ri(max(batch, 1), 128) + 1)
    dst = dict(tokens=ri(cap_tok), positions=ri(cap_tok), new_cache_slots=ri(cap_tok), kv_seq_lens=ri(cap_batch + 1),
               kv_seq_lens_delta=ri(cap_batch), paged_kv_indptr=ri(cap_batch + 1), paged_kv_indices=ri(cap_idx),
               paged_kv_last_page_len=ri(cap_batch))
    return src, dst


@pytest.mark.parametrize("n_tok,padded,batch,n_idx", [(5, 8, 5, 37), (1, 1, 1, 1), (64, 64, 64, 4000), (3, 16, 3, 0), (0, 4, 0, 0),
                                                      (300, 512, 300, 70000)])
def test_decode_metadata_update_bit_exact(n_tok, padded, batch, n_idx, built_lib):
    from xllm_b200 import ops
    cap_tok, cap_batch, cap_idx = max(padded, n_tok) + 7, batch + 5, n_idx + 11
    src, dst = _case(n_tok, padded, batch, n_idx, cap_tok, cap_batch, cap_idx, seed=n_tok * 31 + batch)
    ref = OB.update_llm_decode_metadata({k: v.clone() for k, v in src.items()}, {k: v.clone() for k, v in dst.items()},
                                        n_tok, padded, batch, n_idx)
    dsrc = {k: v.to(DEV) for k, v in src.items()}
    ddst = {k: v.to(DEV) for k, v in dst.items()}
    counters = torch.full((33,), 7, dtype=I32, device=DEV)
    ops.update_llm_decode_metadata(dsrc, ddst, n_tok, padded, batch, n_idx, plan_counters=counters)
    torch.cuda.synchronize()
    for k in ref:
        assert torch.equal(ddst[k].cpu(), ref[k]), k
    assert int(counters.abs().sum()) == 0, "plan counters must be re-zeroed"


def test_graph_replay_serves_a_new_batch_without_host_plan(built_lib):
    """capture {metadata refresh -> paged decode attention} once; replay after changing ONLY the source buffers (new context
    lengths, new pages): the attention output equals the oracle for the new batch - no plan call, no host sync in between
    (the reference re-plans on the host before every replay: cuda_graph_executor_impl.cpp:751-822)."""
    from oracle import ops as O
    from tests.util import assert_close_attention
    from xllm_b200 import ops
    HQ, HKV, D, page, B, max_pages = 28, 4, 128, 16, 3, 300
    BF16 = torch.bfloat16
    g = torch.Generator().manual_seed(5)
    nblocks = B * max_pages + 1
    kc = torch.randn(nblocks, page, HKV, D, generator=g).to(BF16)
    vc = torch.randn(nblocks, page, HKV, D, generator=g).to(BF16)
    q = torch.randn(B, HQ, D, generator=g).to(BF16)

    def batch(kv_lens, seed):
        gg = torch.Generator().manual_seed(seed)
        perm = (torch.randperm(nblocks - 1, generator=gg) + 1).to(I32)
        pages, indptr, last, off = [], [0], [], 0
        for n in kv_lens:
            npg = (n + page - 1) // page
            pages.append(perm[off:off + npg])
            off += npg
            indptr.append(indptr[-1] + npg)
            last.append((n - 1) % page + 1)
        return (torch.tensor(indptr, dtype=I32), torch.cat(pages), torch.tensor(last, dtype=I32),
                torch.tensor([0] + list(torch.tensor(kv_lens).cumsum(0)), dtype=I32))

    cap = max_pages * B
    src = dict(tokens=torch.zeros(B, dtype=I32, device=DEV), positions=torch.zeros(B, dtype=I32, device=DEV),
               new_cache_slots=torch.zeros(B, dtype=I32, device=DEV), kv_seq_lens=torch.zeros(B + 1, dtype=I32, device=DEV),
               paged_kv_indptr=torch.zeros(B + 1, dtype=I32, device=DEV), paged_kv_indices=torch.zeros(cap, dtype=I32, device=DEV),
               paged_kv_last_page_len=torch.ones(B, dtype=I32, device=DEV))
    dst = {k: torch.zeros_like(v) for k, v in src.items()}
    dst["kv_seq_lens_delta"] = torch.zeros(B, dtype=I32, device=DEV)
    plan = ops.DecodePlan(B, HQ, HKV, D, page, max_pages, DEV)
    qd, kcd, vcd = q.to(DEV), kc.to(DEV), vc.to(DEV)
    out = torch.empty(B, HQ, D, dtype=BF16, device=DEV)

    def fill(kv_lens, seed):
        indptr, indices, last, cum = batch(kv_lens, seed)
        src["paged_kv_indptr"].copy_(indptr)
        src["paged_kv_indices"][:indices.numel()].copy_(indices)
        src["paged_kv_last_page_len"].copy_(last)
        src["kv_seq_lens"].copy_(cum)
        return indptr, indices, last

    def launch():
        ops.update_llm_decode_metadata(src, dst, B, B, B, cap, plan_counters=plan.int_ws.view(torch.int32))
        ops.batch_decode(plan, qd, kcd, vcd, dst["paged_kv_indptr"], dst["paged_kv_indices"], dst["paged_kv_last_page_len"],
                         1 / math.sqrt(D), out)

    fill([40, 17, 300], 1)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        launch()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            launch()
    torch.cuda.current_stream().wait_stream(s)
    for kv_lens, seed in (([40, 17, 300], 1), ([4700, 1, 2222], 2), ([16, 4800, 33], 3)):
        indptr, indices, last = fill(kv_lens, seed)
        graph.replay()
        torch.cuda.synchronize()
        qo = torch.arange(B + 1, dtype=I32)
        ref = O.paged_attention(q, kc, vc, qo, indptr, indices, last, 1 / math.sqrt(D), causal=False)
        scale = O.paged_attention(q, kc, vc.abs(), qo, indptr, indices, last, 1 / math.sqrt(D), causal=False)
        assert_close_attention(out, ref, scale, what=f"replay kv_lens={kv_lens}")
