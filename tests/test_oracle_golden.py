"""Pins the oracle against every golden vector the reference tests hold for the path."""
import math

import pytest
import torch

from oracle import batch, layer, ops, seeded

BF16 = torch.bfloat16


# ---- integer page-table metadata: BatchTest.Basic (tests/core/framework/batch/batch_test.cpp:403-546) ---------
def test_batch_basic_golden():
    S = batch.SeqState
    m = batch.build_paged_meta([S([1, 2, 3], 0, 9), S([4, 5, 6, 7], 7, 8), S([8, 9, 10, 11, 12], 15, 16),
                                S([13, 14, 15], 4, 8)], block_size=4)
    assert m.positions == [0, 1, 2, 3, 4, 5, 6, 7, 8, 7, 15, 4, 5, 6, 7]
    assert m.q_cu_seq_lens == [0, 9, 10, 11, 15]
    assert m.kv_cu_seq_lens == [0, 9, 17, 33, 41]
    assert m.new_cache_slots == [4, 5, 6, 7, 8, 9, 10, 11, 12, 23, 47, 56, 57, 58, 59]
    assert m.padded_block_tables() == [1, 2, 3, 0, 0, 4, 5, 6, 7, 0, 8, 9, 10, 11, 12, 13, 14, 15, 0, 0]
    # the paged triplet per batch_input_builder.cpp:790-801
    assert m.paged_kv_indptr == [0, 3, 7, 12, 15]
    assert m.paged_kv_indices == list(range(1, 16))
    assert m.paged_kv_last_page_len == [1, 4, 4, 4]


def test_decode_padding_rows():
    # batch_input_builder.cpp:854-873: padded decode rows use slot 0 / block 0 / last_page_len 1
    S = batch.SeqState
    m = batch.build_paged_meta([S([3, 4], 5, 6)], block_size=4, min_decoding_batch_size=3)
    assert m.new_cache_slots == [17, 0, 0]          # pos 5 -> block 4, offset 1
    assert m.paged_kv_indptr == [0, 2, 3, 4]
    assert m.paged_kv_indices == [3, 4, 0, 0]
    assert m.paged_kv_last_page_len == [2, 1, 1]


# ---- seeded_tensor (tests/core/layers/mlu/tests_utils.cpp:159-274) ------------------------------------------
def test_seeded_tensor_stream():
    assert seeded.fnv1a64("") == 0xCBF29CE484222325
    assert seeded.fnv1a64("a") == 0xAF63DC4C8601EC8C          # published FNV-1a test vector
    # SplitMix64 published vector: seed 1234567 -> first outputs
    s = seeded.splitmix64_stream(1234567, 3).tolist()
    assert s == [6457827717110365317, 3203168211198807973, 9817491932198370423]
    t = seeded.seeded_tensor("k", (2, 3), torch.float32)
    assert t.shape == (2, 3) and float(t.min()) >= 0.0 and float(t.max()) < 1.0


# ---- Qwen2Attention known answers (tests/core/layers/mlu/qwen2_attention_test.cpp:33-393) -------------------
H, NH, NKV, D, BS, NBLK = 1024, 16, 8, 128, 16, 100


def _noise(key, shape, std):
    n = seeded.seeded_tensor(key, shape, BF16)
    return ((n - 0.5) * (math.sqrt(12.0) * std)).to(BF16)      # MakeNoise :118-127 (bf16 tensor, float scalars)


def _weights():
    pre = "qwen2_attention_test."
    def w(name, shape):
        t = seeded.seeded_tensor(pre + name, shape, BF16)
        return (t / torch.sqrt(torch.tensor(float(t.shape[0]), dtype=BF16))).to(BF16)   # :109-112
    q, k, v = w("q_proj.weight", (NH * D, H)), w("k_proj.weight", (NKV * D, H)), w("v_proj.weight", (NKV * D, H))
    qb, kb, vb = w("q_proj.bias", (NH * D,)), w("k_proj.bias", (NKV * D,)), w("v_proj.bias", (NKV * D,))
    o = w("o_proj.weight", (H, NH * D))
    return torch.cat([q, k, v]), torch.cat([qb, kb, vb]), o


def _caches():
    # MLU layout [blocks, heads, block, dim] (:64-71) -> logical NHD [blocks, block, heads, dim]
    k = _noise("qwen2_attention_test.k_cache", (NBLK, NKV, BS, D), 0.01).permute(0, 2, 1, 3).contiguous()
    v = _noise("qwen2_attention_test.v_cache", (NBLK, NKV, BS, D), 0.01).permute(0, 2, 1, 3).contiguous()
    return k, v


def _attn():
    qkv_w, qkv_b, o_w = _weights()
    cs = ops.compute_cos_sin_cache(D, 2048, 1000000.0, BF16)
    return layer.Qwen2AttentionOracle(qkv_w, qkv_b, o_w, NH, NKV, D, cs)


def _blocks(seq_len):
    return (seq_len + BS - 1) // BS + 1                        # GetBlockNum :129-132


def _check(out10, expected):
    got = out10.to(torch.float32)
    exp = torch.tensor(expected, dtype=torch.float32)
    # The stored values come from MLU hardware, checked there with rtol 1e-5 / atol 1e-6
    # (test::verify_precision).  The CPU restatement reproduces them to the printed precision.
    rel = ((got - exp).abs() / exp.abs()).max().item()
    assert torch.allclose(got, exp, rtol=1e-5, atol=1e-6), f"max rel diff {rel:.3e}; got {got.tolist()}"
    print(f"max rel diff vs MLU known answers: {rel:.3e}")
    return rel


def test_qwen2_attention_prefill_kat():
    attn = _attn()
    B, S = 2, 128
    hidden = _noise("qwen2_attention_test.prefill.hidden_states", (B * S, H), 0.02)
    positions = torch.arange(S).repeat(B)
    cu = torch.arange(0, (B + 1) * S, S, dtype=torch.int32)
    nb = _blocks(S)
    slots = torch.tensor([b * nb * BS + i for b in range(B) for i in range(S)], dtype=torch.int32)
    meta = layer.AttnMeta(True, False, cu, cu, slots)
    k_cache, v_cache = _caches()
    out = attn.forward(positions, hidden, meta, k_cache, v_cache)
    _check(out.flatten()[:10], [0.6796875, 0.67578125, 0.6875, 0.65625, 0.6640625, 0.6796875, 0.68359375,
                                0.67578125, 0.6796875, 0.66796875])


def test_qwen2_attention_decode_kat():
    attn = _attn()
    B, S = 4, 256
    kv_len = S + 1
    hidden = _noise("qwen2_attention_test.decode.hidden_states", (B, H), 0.02)
    positions = torch.full((B,), S)
    nb = _blocks(kv_len)
    slots = torch.tensor([b * nb * BS + kv_len - 1 for b in range(B)], dtype=torch.int32)
    q_cu = torch.arange(0, B + 1, dtype=torch.int32)
    kv_cu = torch.arange(0, (B + 1) * kv_len, kv_len, dtype=torch.int32)
    # block table b*nb + i  ->  paged triplet covering ceil(kv_len/BS) pages
    npg = (kv_len + BS - 1) // BS
    indptr = torch.arange(0, (B + 1) * npg, npg, dtype=torch.int32)
    indices = torch.tensor([b * nb + i for b in range(B) for i in range(npg)], dtype=torch.int32)
    last = torch.full((B,), kv_len - (npg - 1) * BS, dtype=torch.int32)
    meta = layer.AttnMeta(False, False, q_cu, kv_cu, slots, indptr, indices, last)
    k_cache, v_cache = _caches()
    out = attn.forward(positions, hidden, meta, k_cache, v_cache)
    _check(out.flatten()[:10], [0.0005264282, 0.0008239746, 0.0005722046, 0.0006027222, 0.000831604, 0.0004405975,
                                0.001037598, 0.001083374, 0.000289917, 0.0007820129])


def test_qwen2_attention_mixed_prefill_kat():
    attn = _attn()
    lens = [32, 64, 128]
    total = sum(lens)
    hidden = _noise("qwen2_attention_test.mix.hidden_states", (total, H), 0.02)
    positions = torch.cat([torch.arange(n) for n in lens])
    cu = torch.tensor([0, 32, 96, 224], dtype=torch.int32)
    meta = layer.AttnMeta(True, False, cu, cu, torch.arange(total, dtype=torch.int32))
    k_cache, v_cache = _caches()
    out = attn.forward(positions, hidden, meta, k_cache, v_cache)
    _check(out.flatten()[:10], [0.07763672, 0.08349609, 0.08496094, 0.08349609, 0.07958984, 0.08740234, 0.09130859,
                                0.08398438, 0.08642578, 0.07958984])


def test_update_llm_decode_metadata_restatement():
    """hand-checked case of llm_decode_metadata_update_kernel (llm_decode_metadata_update.cu:29-62): 2 live requests in a
    graph captured for 4, persistent buffers larger than the step."""
    import numpy as np
    from oracle import batch as OB
    i32 = lambda *v: np.array(v, dtype=np.int32)
    src = dict(tokens=i32(11, 12), positions=i32(7, 3), new_cache_slots=i32(39, 20), kv_seq_lens=i32(0, 8, 12),
               paged_kv_indptr=i32(0, 2, 3), paged_kv_indices=i32(5, 9, 4), paged_kv_last_page_len=i32(4, 4))
    dst = dict(tokens=i32(9, 9, 9, 9, 9), positions=i32(9, 9, 9, 9, 9), new_cache_slots=i32(9, 9, 9, 9, 9),
               kv_seq_lens=i32(9, 9, 9, 9, 9), kv_seq_lens_delta=i32(9, 9, 9, 9), paged_kv_indptr=i32(9, 9, 9, 9, 9),
               paged_kv_indices=i32(9, 9, 9, 9, 9, 9), paged_kv_last_page_len=i32(9, 9, 9, 9))
    OB.update_llm_decode_metadata(src, dst, actual_num_tokens=2, padded_num_tokens=4, actual_batch_size=2, actual_indices_size=3)
    assert dst["tokens"].tolist() == [11, 12, 0, 0, 9]            # padding rows zeroed, tail untouched
    assert dst["positions"].tolist() == [7, 3, 9, 9, 9]           # positions of padding rows are NOT written
    assert dst["new_cache_slots"].tolist() == [39, 20, 0, 0, 9]
    assert dst["kv_seq_lens"].tolist() == [0, 8, 12, 9, 9]
    assert dst["kv_seq_lens_delta"].tolist() == [8, 4, 9, 9]
    assert dst["paged_kv_indptr"].tolist() == [0, 2, 3, 9, 9]
    assert dst["paged_kv_indices"].tolist() == [5, 9, 4, 9, 9, 9]
    assert dst["paged_kv_last_page_len"].tolist() == [4, 4, 9, 9]
