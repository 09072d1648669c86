"""SURVEY 8(d) cfg1 - a prompt is prefilled (one shot and in chunks) through
Qwen2PrefillRunner, then greedy-decoded through Qwen2DecodeRunner on the SAME paged KV cache, against the CPU oracle:
same tokens, last-token logits within the whole-step tolerance of tests/test_gpu_model.py.
"""
import math
import os

import pytest
import torch

from oracle import batch as OB
from tests import model_parity as MP

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
DEV = "cuda"


def _rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("quant", ["bf16", "w4a16"])
@pytest.mark.parametrize("chunks", [1, 3])
def test_prefill_then_greedy_decode_matches_oracle(quant, chunks, built_lib):
    from xllm_b200.qwen2 import Qwen2Config, Qwen2DecodeRunner
    from xllm_b200.qwen2_prefill import Qwen2PrefillRunner
    cfg = Qwen2Config.qwen2_0_5b(num_layers=2, vocab_size=4096, block_size=16, max_position_embeddings=512, quant=quant,
                                 tie_word_embeddings=False)
    W, _, _, _ = MP.build_case(cfg, 1, [1], seed=11)
    g = torch.Generator().manual_seed(5)
    prompt_len, n_decode = 128, 4
    bs = cfg.block_size
    total = prompt_len + n_decode
    nblk = (total + bs - 1) // bs
    nblocks = nblk + 4
    blocks = (torch.randperm(nblocks - 1, generator=g) + 1)[:nblk].tolist()
    prompt = torch.randint(0, cfg.vocab_size, (prompt_len,), generator=g).tolist()

    # ---- oracle: one-shot prefill + greedy decode -----------------------------------------------------------------
    mk = lambda: [torch.zeros(nblocks, bs, cfg.n_kv_heads, cfg.head_dim, dtype=BF16) for _ in range(cfg.num_layers)]
    kc_o, vc_o = mk(), mk()
    used = lambda n: blocks[: (n + bs - 1) // bs]
    meta = OB.build_paged_meta([OB.SeqState(used(prompt_len), 0, prompt_len)], bs)
    ref_logits = MP.oracle_prefill(cfg, W, kc_o, vc_o, prompt, meta, chunked=False)
    ref_tokens = [int(ref_logits.float().argmax(-1))]
    for i in range(n_decode - 1):
        n = prompt_len + i + 1
        m = OB.build_paged_meta([OB.SeqState(used(n), n - 1, n)], bs)
        step = dict(tokens=[ref_tokens[-1]], positions=m.positions, slots=m.new_cache_slots, indptr=m.paged_kv_indptr,
                    indices=m.paged_kv_indices, last=m.paged_kv_last_page_len, nblocks=nblocks)
        _, nxt = MP.oracle_step(cfg, W, kc_o, vc_o, step)
        ref_tokens.append(int(nxt[0]))

    # ---- GPU: prefill (1 or 3 chunks) + decode on the same cache -------------------------------------------------
    runner = Qwen2DecodeRunner(cfg, MP.upload(cfg, W), max_batch=1, max_ctx=nblk * bs, device=DEV, num_blocks=nblocks)
    pre = Qwen2PrefillRunner.from_decode_runner(runner)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    bounds = [prompt_len * i // chunks for i in range(chunks + 1)]
    logits = tokens = None
    for ci in range(chunks):
        a, b = bounds[ci], bounds[ci + 1]
        m = OB.build_paged_meta([OB.SeqState(used(b), a, b)], bs)
        logits, tokens = pre.forward(i32(prompt[a:b]), torch.tensor(m.positions, dtype=torch.int64, device=DEV),
                                     i32(m.new_cache_slots), i32(m.q_cu_seq_lens), i32(m.kv_cu_seq_lens),
                                     i32(m.paged_kv_indptr), i32(m.paged_kv_indices), i32(m.paged_kv_last_page_len),
                                     chunked=(a > 0), max_qo_len=b - a)
    torch.cuda.synchronize()
    assert _rel_l2(logits, ref_logits) <= 2e-2, "prefill logits"
    got = [int(tokens[0])]
    top2 = ref_logits.float().topk(2, -1).values[0]
    if float(top2[0] - top2[1]) > 0.05:                      # a near-tie may legitimately flip under bf16 rounding
        assert got[0] == ref_tokens[0]
    for i in range(n_decode - 1):
        n = prompt_len + i + 1
        m = OB.build_paged_meta([OB.SeqState(used(n), n - 1, n)], bs)
        runner.set_inputs_host([ref_tokens[i]], m.positions, m.new_cache_slots, m.paged_kv_indptr, m.paged_kv_indices,
                               m.paged_kv_last_page_len)       # teacher-forced on the oracle's token: isolates each step
        got.append(int(runner.step()[0]))
    assert got[1:] == ref_tokens[1:] or sum(x != y for x, y in zip(got, ref_tokens)) <= 1, (got, ref_tokens)


def test_cfg0_qwen2_0_5b_full_model_prompt128_greedy16(built_lib):
    """BASELINE configs[0] in full: Qwen2-0.5B bf16 (24 layers, hidden 896, 14/2 heads of 64, vocab 151936, tied
    embeddings), batch 1, prompt 128 prefilled in one shot, then 16 greedy decode steps on the same paged KV cache
    (block_size 128: the prompt fills page 0 exactly, decode continues on a second, non-adjacent page) - the composition
    of LlmModelImplBase::forward (xllm/models/llm/llm_model_base.h:60-131) against the CPU oracle.

    Two bf16 pipelines that are both correct drift by ~1 bf16 ulp per op, so "token-exact" is asserted wherever the
    oracle's own top-2 margin exceeds that drift (|logit1 - logit2| > max(2 % of |logit1|, 4 bf16 ulps of logit1 - the
    logits themselves are bf16)); every GPU token must in any case be a near-argmax of the oracle's logits.  Decode steps are teacher-forced on the oracle's token so each step is
    checked in isolation."""
    from xllm_b200.qwen2 import Qwen2Config, Qwen2DecodeRunner
    from xllm_b200.qwen2_prefill import Qwen2PrefillRunner
    cfg = Qwen2Config.qwen2_0_5b()
    assert (cfg.num_layers, cfg.hidden_size, cfg.vocab_size, cfg.block_size) == (24, 896, 151936, 128)
    # weights N(0, 0.02^2) like a real initialisation: with the helper's default 0.05 the 24-layer random stack has a
    # per-layer gain > 1 and amplifies bf16 rounding noise chaotically (measured 5e-2 logits rel-L2 between two correct
    # bf16 pipelines, argmax flips at 3-ulp margins)
    W, _, _, _ = MP.build_case(cfg, 1, [1], seed=2026, w_std=0.02)
    g = torch.Generator().manual_seed(2026)
    prompt_len, n_decode = 128, 16
    bs = cfg.block_size
    nblk = (prompt_len + n_decode + bs - 1) // bs
    nblocks = nblk + 5
    blocks = (torch.randperm(nblocks - 1, generator=g) + 1)[:nblk].tolist()
    prompt = torch.randint(0, cfg.vocab_size, (prompt_len,), generator=g).tolist()
    used = lambda n: blocks[: (n + bs - 1) // bs]
    mk = lambda: [torch.zeros(nblocks, bs, cfg.n_kv_heads, cfg.head_dim, dtype=BF16) for _ in range(cfg.num_layers)]
    kc_o, vc_o = mk(), mk()
    meta0 = OB.build_paged_meta([OB.SeqState(used(prompt_len), 0, prompt_len)], bs)
    ref_logits = [MP.oracle_prefill(cfg, W, kc_o, vc_o, prompt, meta0, chunked=False)]
    ref_tokens = [int(ref_logits[0].float().argmax(-1))]
    metas = []
    for i in range(n_decode - 1):
        n = prompt_len + i + 1
        m = OB.build_paged_meta([OB.SeqState(used(n), n - 1, n)], bs)
        metas.append(m)
        step = dict(tokens=[ref_tokens[-1]], positions=m.positions, slots=m.new_cache_slots, indptr=m.paged_kv_indptr,
                    indices=m.paged_kv_indices, last=m.paged_kv_last_page_len, nblocks=nblocks)
        lg, nxt = MP.oracle_step(cfg, W, kc_o, vc_o, step)
        ref_logits.append(lg)
        ref_tokens.append(int(nxt[0]))

    runner = Qwen2DecodeRunner(cfg, MP.upload(cfg, W), max_batch=1, max_ctx=nblk * bs, device=DEV, num_blocks=nblocks)
    pre = Qwen2PrefillRunner.from_decode_runner(runner)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    logits, tokens = pre.forward(i32(prompt), torch.tensor(meta0.positions, dtype=torch.int64, device=DEV),
                                 i32(meta0.new_cache_slots), i32(meta0.q_cu_seq_lens), i32(meta0.kv_cu_seq_lens),
                                 i32(meta0.paged_kv_indptr), i32(meta0.paged_kv_indices), i32(meta0.paged_kv_last_page_len),
                                 chunked=False, max_qo_len=prompt_len)
    torch.cuda.synchronize()
    got_tokens, got_logits = [int(tokens[0])], [logits.cpu()]
    # the decode steps replay ONE captured CUDA graph (capture happens at the first step's inputs)
    for i, m in enumerate(metas):
        runner.set_inputs_host([ref_tokens[i]], m.positions, m.new_cache_slots, m.paged_kv_indptr, m.paged_kv_indices,
                               m.paged_kv_last_page_len)
        if i == 1:
            # capture() runs the step twice (warm-up + capture) on the current inputs: that re-writes the same KV row
            # with the same values, which is idempotent
            runner.step()
            runner.capture()
        got_tokens.append(int(runner.step()[0]))
        got_logits.append(runner.logits[:1].cpu().clone())
    worst, exact, decided, report, bad = 0.0, 0, 0, [], []
    for i in range(n_decode):
        rl = ref_logits[i].float()[0]
        l2 = _rel_l2(got_logits[i], ref_logits[i])
        worst = max(worst, l2)
        top2 = rl.topk(2).values
        margin = float(top2[0] - top2[1])
        top = abs(float(top2[0]))
        tol = max(0.02 * top, 4 * 2.0 ** (math.floor(math.log2(max(top, 1e-30))) - 7))
        same = got_tokens[i] == ref_tokens[i]
        exact += int(same)
        near = float(rl[got_tokens[i]]) >= float(top2[0]) - tol
        report.append(f"step {i}: rel_l2 {l2:.2e} top {float(top2[0]):.2f} margin {margin:.2f} tol {tol:.2f} "
                      f"got {got_tokens[i]} ref {ref_tokens[i]} {'ok' if same else ('near-tie' if near else 'MISMATCH')}")
        if margin > tol:
            decided += 1
            if not same:
                bad.append(i)
        if not near:
            bad.append(i)
    from tests.util import REL_L2_LOG
    REL_L2_LOG.append((f"cfg0 Qwen2-0.5B full model: worst logits rel-L2 over {n_decode} steps; {exact}/{n_decode} tokens "
                       f"exact, {decided} steps with a decisive oracle margin", worst))
    assert not bad and worst <= 3e-2 and exact >= n_decode - 2, "\n".join(report)
