"""CPU check of the prefill / chunked-prefill COMPOSITION (xllm_b200/qwen2_prefill.py): the runner's sequence of op calls,
buffer aliasing, strided q/k/v views, slot / page metadata (oracle.batch = the reference's batch builder) and
last-token selection, with the library ops replaced by an adapter that executes the oracle's arithmetic in place.
Against the oracle's own layer composition the result must be bit-identical; chunked prefill must equal one-shot
prefill exactly (the cache holds the same bf16 K/V the ragged path reads).  The kernels themselves are covered by the
-m gpu tests; nothing here touches the CUDA library."""
import torch

from oracle import batch as OB
from oracle import layer as OL
from oracle import ops as O
from tests import model_parity as MP
from xllm_b200 import qwen2 as Q2
from xllm_b200.qwen2 import Qwen2Config
from xllm_b200.qwen2_prefill import Qwen2PrefillRunner

BF16 = torch.bfloat16


class OracleOps:
    """Same call signatures and in-place conventions as xllm_b200.ops, executed by the oracle on CPU tensors."""

    @staticmethod
    def embedding(out, token_ids, table):
        out.copy_(table[token_ids.long()])

    @staticmethod
    def rms_norm(output, input, weight, eps):
        output.copy_(O.rms_norm(input, weight, eps))

    @staticmethod
    def fused_add_rms_norm(input, residual, weight, eps):
        h, r = O.fused_add_rms_norm(input, residual, weight, eps)
        input.copy_(h)
        residual.copy_(r)

    @staticmethod
    def rope_and_cache(positions, query, key, value, cos_sin_cache, slot_ids, key_cache, value_cache, is_neox=True):
        T, D, nkv = positions.numel(), key_cache.size(-1), key_cache.size(-2)
        q3, k3 = query.reshape(T, -1, D), key.reshape(T, nkv, D)
        q2, k2 = O.rotary_embedding(positions, q3, k3, cos_sin_cache, is_neox=is_neox)
        query.copy_(q2.reshape(T, -1))
        key.copy_(k2.reshape(T, -1))
        O.reshape_paged_cache(slot_ids, key.reshape(T, nkv, D), value.reshape(T, nkv, D), key_cache, value_cache)

    @staticmethod
    def batch_prefill(query, key, value, q_cu, kv_cu, sm_scale, output, output_lse=None, max_qo_len=None, causal=True):
        assert max_qo_len >= int((q_cu[1:] - q_cu[:-1]).max())
        output.copy_(O.ragged_prefill_attention(query, key, value, q_cu, kv_cu, sm_scale, causal=causal))

    @staticmethod
    def batch_chunked_prefill(query, k_cache, v_cache, indptr, indices, last, sm_scale, output, output_lse=None,
                              qo_indptr=None, causal=True, max_qo_len=None):
        output.copy_(O.paged_attention(query, k_cache, v_cache, qo_indptr, indptr, indices, last, sm_scale, causal=causal))

    @staticmethod
    def act_and_mul(out, input, act_mode):
        out.copy_(O.act_and_mul(input, act_mode))

    @staticmethod
    def matmul(a, b, bias=None, out=None):
        out.copy_(O.linear(a, b, bias))
        return out

    @staticmethod
    def argmax(out, logits):
        out.copy_(logits.float().argmax(-1).to(torch.int32))


def _tiny_cfg():
    return Qwen2Config(hidden_size=64, num_layers=2, n_heads=4, n_kv_heads=2, head_dim=16, intermediate_size=96,
                       vocab_size=101, max_position_embeddings=64, block_size=4, quant="bf16", name="tiny")


_oracle_forward = MP.oracle_prefill


def _runner_forward(cfg, W, kcs, vcs, tokens, meta: OB.PagedMeta, chunked, monkeypatch):
    monkeypatch.setattr(Q2, "ops", OracleOps)            # Linear.forward and the runner both resolve ops through qwen2
    w = MP.upload(cfg, W, device="cpu")
    cs = Q2.make_cos_sin_cache(cfg, "cpu")
    r = Qwen2PrefillRunner(cfg, w, kcs, vcs, cs, device="cpu")
    i32 = lambda v: torch.tensor(v, dtype=torch.int32)
    return r.forward(i32(tokens), torch.tensor(meta.positions, dtype=torch.int64), i32(meta.new_cache_slots),
                     i32(meta.q_cu_seq_lens), i32(meta.kv_cu_seq_lens), i32(meta.paged_kv_indptr), i32(meta.paged_kv_indices),
                     i32(meta.paged_kv_last_page_len), chunked=chunked)


def _fresh_caches(cfg, nblocks):
    mk = lambda: [torch.zeros(nblocks, cfg.block_size, cfg.n_kv_heads, cfg.head_dim, dtype=BF16) for _ in range(cfg.num_layers)]
    return mk(), mk()


def test_prefill_composition_matches_oracle_and_chunking_is_exact(monkeypatch):
    cfg = _tiny_cfg()
    W, _, _, _ = MP.build_case(cfg, 1, [1], seed=7)
    g = torch.Generator().manual_seed(3)
    lens = [5, 9, 1]
    blocks = [[3, 1], [6, 2, 5], [4]]                      # scattered physical blocks, block 0 reserved
    toks = [torch.randint(0, cfg.vocab_size, (n,), generator=g).tolist() for n in lens]
    flat = [t for s in toks for t in s]
    nblocks = 8

    # ---- one-shot prefill ---------------------------------------------------------------------------------------
    meta = OB.build_paged_meta([OB.SeqState(b, 0, n) for b, n in zip(blocks, lens)], cfg.block_size)
    kc_o, vc_o = _fresh_caches(cfg, nblocks)
    ref = _oracle_forward(cfg, W, kc_o, vc_o, flat, meta, chunked=False)
    kc_r, vc_r = _fresh_caches(cfg, nblocks)
    logits, tokens = _runner_forward(cfg, W, kc_r, vc_r, flat, meta, False, monkeypatch)
    assert torch.equal(logits, ref), "runner composition differs from the oracle composition"
    assert torch.equal(tokens.long(), ref.float().argmax(-1))
    for a, b in zip(kc_r + vc_r, kc_o + vc_o):
        assert torch.equal(a, b), "KV cache contents differ"

    # ---- the same prompts in two chunks: [first 3 | rest] (sequence 2 has a single token: it finishes in chunk 1) ----
    cut = [3, 3, 1]
    m1 = OB.build_paged_meta([OB.SeqState(b[: (c + cfg.block_size - 1) // cfg.block_size], 0, c) for b, c in zip(blocks, cut)],
                             cfg.block_size)
    kc_c, vc_c = _fresh_caches(cfg, nblocks)
    first = [t for s, c in zip(toks, cut) for t in s[:c]]
    _runner_forward(cfg, W, kc_c, vc_c, first, m1, False, monkeypatch)
    rest_seqs = [(b, c, n) for b, c, n in zip(blocks, cut, lens) if n > c]
    m2 = OB.build_paged_meta([OB.SeqState(b, c, n) for b, c, n in rest_seqs], cfg.block_size)
    second = [t for s, c, n in zip(toks, cut, lens) if n > c for t in s[c:]]
    logits2, tokens2 = _runner_forward(cfg, W, kc_c, vc_c, second, m2, True, monkeypatch)
    assert torch.equal(logits2, ref[:2]), "chunked prefill must reproduce one-shot prefill exactly"
    assert torch.equal(tokens2, tokens[:2])
    for a, b in zip(kc_c + vc_c, kc_o + vc_o):
        assert torch.equal(a, b), "chunked prefill left a different KV cache"
