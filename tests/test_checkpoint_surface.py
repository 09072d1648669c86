"""CPU tests of the checkpoint surface: AutoAWQ / AutoGPTQ tensor conventions -> logical form ->
kernel layout.  The packers below restate the two libraries' published packing independently of the converters."""
import pytest
import torch

from oracle import quant as OQ
from xllm_b200 import quant

AWQ_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]


def _logical(N=64, K=256, gs=128, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = torch.randint(0, 16, (N, K), dtype=torch.uint8, generator=g)
    z = torch.randint(1, 16, (N, K // gs), dtype=torch.uint8, generator=g)
    s = (torch.rand(N, K // gs, generator=g) * 0.02 + 0.001).to(torch.float16)
    return q, z, s


def _pack_awq(vals_kn):                     # [K, N] -> int32 [K, N/8], AutoAWQ order_map
    K, N = vals_kn.shape
    v = vals_kn.to(torch.int64).view(K, N // 8, 8)
    word = torch.zeros(K, N // 8, dtype=torch.int64)
    for i, col in enumerate(AWQ_ORDER):
        word |= v[:, :, col] << (4 * i)
    return word.to(torch.int32)             # wraps bit 31 into the sign like the real checkpoints


def _pack_seq_lastdim(vals):                # [..., 8C] -> int32 [..., C], low nibble first
    v = vals.to(torch.int64).view(*vals.shape[:-1], vals.shape[-1] // 8, 8)
    word = torch.zeros(v.shape[:-1], dtype=torch.int64)
    for i in range(8):
        word |= v[..., i] << (4 * i)
    return word.to(torch.int32)


def test_from_awq_roundtrip():
    q, z, s = _logical()
    qw = _pack_awq(q.t().contiguous())
    qz = _pack_awq(z.t().contiguous())
    q2, s2, z2 = quant.from_awq(qw, qz, s.t().contiguous(), 128)
    assert torch.equal(q2, q) and torch.equal(z2, z) and torch.equal(s2, s.to(torch.bfloat16))


def test_from_gptq_roundtrip_and_desc_act_rejected():
    q, z, s = _logical(seed=1)
    N, K = q.shape
    qw = _pack_seq_lastdim(q).t().contiguous()                  # [K/8, N]
    qz = _pack_seq_lastdim((z.t().to(torch.int16) - 1).contiguous())   # stores z - 1, packed along N
    g_idx = torch.arange(K, dtype=torch.int32) // 128
    q2, s2, z2 = quant.from_gptq(qw, qz, s.t().contiguous(), g_idx, 128)
    assert torch.equal(q2, q) and torch.equal(z2, z) and torch.equal(s2, s.to(torch.bfloat16))
    with pytest.raises(ValueError):
        quant.from_gptq(qw, qz, s.t().contiguous(), g_idx.flip(0), 128)


def test_checkpoint_to_kernel_layout_matches_spec():
    """AWQ tensors -> logical -> kernel layout; the packed words decode (by the documented nibble positions) back to the
    oracle's dequantised weight."""
    q, z, s = _logical(N=32, K=128, seed=2)
    q2, s2, z2 = quant.from_awq(_pack_awq(q.t().contiguous()), _pack_awq(z.t().contiguous()), s.t().contiguous(), 128)
    qw, meta = quant.pack_w4(q2, s2, z2, 128)
    wd = OQ.dequantize(q2, s2, z2, 128).float()
    # decode tile (nt=1, kt=1), lane (g=3, t=2), word j=1: rows 16+3 / 16+11, k = 64 + 32 + 4 + {0..3}
    w = int(qw[1, 1, 3 * 4 + 2, 1]) & 0xFFFFFFFF
    nib = [(w >> (4 * i)) & 15 for i in range(8)]
    k0 = 64 + 16 * 2 + 4
    assert [nib[0], nib[4], nib[2], nib[6]] == q2[19, k0:k0 + 4].tolist()
    assert [nib[1], nib[5], nib[3], nib[7]] == q2[27, k0:k0 + 4].tolist()
    m = int(meta[0, 19]) & 0xFFFFFFFF
    sc = torch.tensor([m & 0xFFFF], dtype=torch.int32).to(torch.int16).view(torch.bfloat16).float().item()
    zb = torch.tensor([m >> 16], dtype=torch.int32).to(torch.int16).view(torch.bfloat16).float().item()
    assert zb == 128 + int(z2[19, 0])
    ref = torch.tensor((nib[0] - int(z2[19, 0])) * sc).to(torch.bfloat16).float().item()
    assert ref == wd[19, k0].item()


def test_llama3_rope_scaling_matches_the_published_rule():
    """xllm_b200.qwen2.llama3_scale_inv_freq (vectorised) against the element-by-element restatement in the oracle;
    the scaled table changes only the low frequencies (Llama-3.1: factor 8, low 1, high 4, original 8192)."""
    import torch
    from oracle import ops as O
    from xllm_b200.qwen2 import Qwen2Config, llama3_scale_inv_freq, make_cos_sin_cache
    inv = O.compute_inv_freq(128, 500000.0)
    got = llama3_scale_inv_freq(inv, 8.0, 1.0, 4.0, 8192)
    ref = O.llama3_inv_freq(inv, 8.0, 1.0, 4.0, 8192)
    assert torch.allclose(got, ref, rtol=1e-6, atol=0)
    assert torch.equal(got[:20], inv[:20]) and torch.allclose(got[-5:], inv[-5:] / 8.0)
    plain = make_cos_sin_cache(Qwen2Config.llama3_70b(max_position_embeddings=256), "cpu")
    scaled = make_cos_sin_cache(Qwen2Config.llama3_70b(max_position_embeddings=256, rope_scaling=dict(
        rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=8192)), "cpu")
    assert torch.equal(plain[:, :20], scaled[:, :20]) and not torch.equal(plain[:, 40:64], scaled[:, 40:64])


def test_llama3_rope_scaling_matches_transformers():
    """the same frequencies as the public implementation the checkpoints are trained with (transformers' "llama3" rope init):
    bit-identical inv_freq for Llama-3.1-70B's parameters"""
    pytest = __import__("pytest")
    tf = pytest.importorskip("transformers")
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    from xllm_b200.qwen2 import llama3_scale_inv_freq
    rs = dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=8192)
    try:
        cfg = tf.LlamaConfig(hidden_size=8192, num_attention_heads=64, rope_theta=500000.0, max_position_embeddings=131072, rope_scaling=rs)
        inv, att = ROPE_INIT_FUNCTIONS["llama3"](cfg, "cpu")
    except Exception as e:                                                # config surface differs between transformers versions
        pytest.skip(f"transformers rope init not callable here: {e}")
    sl = torch.arange(0, 128, 2, dtype=torch.float32)
    base = 1.0 / torch.pow(torch.tensor(500000.0), sl / 128.0)
    assert att == 1.0 and torch.equal(inv, llama3_scale_inv_freq(base, 8.0, 1.0, 4.0, 8192))


def test_awq_gptq_unpackers_match_vllm_pack_definitions():
    """xllm_b200.quant.from_awq / from_gptq against an independent public definition of the two checkpoint layouts: vLLM's
    awq_pack / gptq_pack / pack_cols (model_executor/layers/quantization/utils/quant_utils.py) applied to random 4-bit weights and
    zero points must round-trip through our loaders."""
    pytest = __import__("pytest")
    try:
        from vllm.model_executor.layers.quantization.utils import quant_utils as qu
    except Exception as e:
        pytest.skip(f"vllm quant utils not importable here: {e}")
    from xllm_b200 import quant
    g = torch.Generator().manual_seed(4)
    K, N, gs = 256, 64, 128
    q = torch.randint(0, 16, (K, N), generator=g, dtype=torch.int32)            # logical [K, N] nibbles (checkpoint orientation)
    z = torch.randint(0, 16, (K // gs, N), generator=g, dtype=torch.int32)
    s = (torch.rand(K // gs, N, generator=g) * 0.01 + 0.001).to(torch.float16)
    # AutoAWQ GEMM layout: qweight [K, N/8] and qzeros [K/g, N/8], nibbles interleaved 0 2 4 6 1 3 5 7
    q2, s2, z2 = quant.from_awq(qu.awq_pack(q, 4, K, N), qu.awq_pack(z, 4, K // gs, N), s, gs)
    assert torch.equal(q2.to(torch.int32), q.t()) and torch.equal(z2.to(torch.int32), z.t())
    assert torch.equal(s2, s.to(torch.bfloat16).t())
    # AutoGPTQ layout: qweight [K/8, N] packed along K, qzeros [K/g, N/8] packed along N in plain order, stored minus one
    zm1 = (z - 1).clamp(min=0)
    q3, s3, z3 = quant.from_gptq(qu.gptq_pack(q, 4, K, N), qu.pack_cols(zm1, 4, K // gs, N), s, None, gs)
    assert torch.equal(q3.to(torch.int32), q.t()) and torch.equal(z3.to(torch.int32), (zm1 + 1).t())
