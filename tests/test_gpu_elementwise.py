"""GPU parity: elementwise ops through the C ABI vs the oracle (same seeded inputs)."""
import pytest
import torch

from oracle import ops as O
from tests.util import assert_close_bf16

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
DEV = "cuda"


def _g(seed=2026):
    return torch.Generator().manual_seed(seed)


def _randn(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).to(BF16)


@pytest.mark.parametrize("T,H", [(1, 3584), (7, 896), (300, 3584), (3, 8192), (2, 20480), (5, 100)])
def test_rms_norm(T, H, built_lib):
    from xllm_b200 import ops
    g = _g()
    x, w = _randn((T, H), g), (1 + 0.1 * torch.randn(H, generator=g)).to(BF16)
    ref = O.rms_norm(x, w, 1e-6)
    out = torch.empty(T, H, dtype=BF16, device=DEV)
    ops.rms_norm(out, x.to(DEV), w.to(DEV), 1e-6)
    assert_close_bf16(out, ref, what=f"rms_norm {T}x{H}")


def test_rms_norm_strided_rows(built_lib):
    from xllm_b200 import ops
    g = _g(1)
    big = _randn((9, 4608), g)
    x = big[:, 512:512 + 1024]                       # strided view like q slices of qkv
    w = torch.ones(1024).to(BF16)
    ref = O.rms_norm(x, w, 1e-5)
    out = torch.empty(9, 1024, dtype=BF16, device=DEV)
    xd = big.to(DEV)[:, 512:512 + 1024]
    ops.rms_norm(out, xd, w.to(DEV), 1e-5)
    assert_close_bf16(out, ref, what="rms_norm strided")


@pytest.mark.parametrize("T,H", [(1, 3584), (64, 3584), (300, 896), (2, 20480), (3, 100)])
def test_fused_add_rms_norm(T, H, built_lib):
    from xllm_b200 import ops
    g = _g(2)
    x, r, w = _randn((T, H), g), _randn((T, H), g, 3.0), (1 + 0.1 * torch.randn(H, generator=g)).to(BF16)
    ref, ref_res = O.fused_add_rms_norm(x, r, w, 1e-6)
    xd, rd = x.to(DEV), r.to(DEV)
    ops.fused_add_rms_norm(xd, rd, w.to(DEV), 1e-6)
    assert torch.equal(rd.cpu(), ref_res), "residual update must be bit-exact (single bf16 add)"
    assert_close_bf16(xd, ref, what=f"fused_add_rms_norm {T}x{H}")


def _fp8_close(got, ref, what):
    g, r = got.cpu().to(torch.float32), ref.to(torch.float32)
    # e4m3 has 3 mantissa bits: neighbouring codes differ by 2^-3 relative; allow 1 code on <1% of elements
    # (fp32 variance order noise moving a value across a rounding boundary)
    diff = (g - r).abs()
    tol = r.abs().clamp_min(2.0 ** -6) * 2.0 ** -3 + 1e-12
    assert (diff <= tol).all(), f"{what}: more than one e4m3 code apart"
    frac = (diff > 0).float().mean().item()
    assert frac < 0.01, f"{what}: {frac:.4f} of elements differ"


@pytest.mark.parametrize("T,H", [(1, 3584), (33, 1024), (4, 100)])
def test_rms_norm_fp8(T, H, built_lib):
    from xllm_b200 import ops
    g = _g(3)
    x, r, w = _randn((T, H), g), _randn((T, H), g), (1 + 0.1 * torch.randn(H, generator=g)).to(BF16)
    scale = torch.tensor([0.02], dtype=torch.float32)
    ref = O.rms_norm_static_fp8_quant(x, w, scale, 1e-6)
    out = torch.empty(T, H, dtype=torch.float8_e4m3fn, device=DEV)
    ops.rms_norm_static_fp8_quant(out, x.to(DEV), w.to(DEV), scale.to(DEV), 1e-6)
    _fp8_close(out, ref, "rms_norm_static_fp8_quant")
    vec = H % 8 == 0
    ref2, ref_res = O.fused_add_rms_norm_static_fp8_quant(x, r, w, scale, 1e-6, vectorized=vec)
    out2 = torch.empty(T, H, dtype=torch.float8_e4m3fn, device=DEV)
    xd, rd = x.to(DEV), r.to(DEV)
    ops.fused_add_rms_norm_static_fp8_quant(out2, xd, rd, w.to(DEV), scale.to(DEV), 1e-6)
    assert torch.equal(rd.cpu(), ref_res)
    _fp8_close(out2, ref2, "fused_add_rms_norm_static_fp8_quant")


def test_fp8_quant_static_and_dynamic(built_lib):
    from xllm_b200 import ops
    g = _g(4)
    x = _randn((17, 1024), g, 5.0)
    x[0, 0] = 3000.0                                   # saturates at 448
    scale = torch.tensor([0.5], dtype=torch.float32)
    ref = O.static_scaled_fp8_quant(x, scale)
    out = torch.empty(17, 1024, dtype=torch.float8_e4m3fn, device=DEV)
    ops.static_scaled_fp8_quant(out, x.to(DEV), scale.to(DEV))
    assert torch.equal(out.cpu().view(torch.uint8), ref.view(torch.uint8)), "static fp8 quant must be bit-exact"
    ref_q, ref_s = O.fp8_scaled_quantize(x)
    q, s = ops.fp8_scaled_quantize(x.to(DEV))
    assert torch.equal(s.cpu(), ref_s), "dynamic scale must be bit-exact (bf16 amax/448)"
    assert torch.equal(q.cpu().view(torch.uint8), ref_q.view(torch.uint8))


@pytest.mark.parametrize("neox", [True, False])
@pytest.mark.parametrize("T,HQ,HK,D", [(1, 28, 4, 128), (37, 14, 2, 64), (5, 8, 8, 128)])
def test_rotary_embedding_bit_exact(T, HQ, HK, D, neox, built_lib):
    from xllm_b200 import ops
    g = _g(5)
    qkv = _randn((T, (HQ + 2 * HK) * D), g)
    pos = torch.randint(0, 4096, (T,), generator=g)
    cs = O.compute_cos_sin_cache(D, 4096, 1000000, BF16)
    q = qkv[:, :HQ * D].reshape(T, HQ, D)
    k = qkv[:, HQ * D:(HQ + HK) * D].reshape(T, HK, D)
    rq, rk = O.rotary_embedding(pos, q, k, cs, is_neox=neox)
    qkv_d = qkv.to(DEV)
    qd, kd = qkv_d[:, :HQ * D], qkv_d[:, HQ * D:(HQ + HK) * D]   # strided views, as the reference passes them
    ops.rotary_embedding(pos.to(DEV), qd, kd, cs.to(DEV), neox)
    assert torch.equal(qd.cpu().reshape(T, HQ, D), rq), "RoPE(q) must be bit-exact"
    assert torch.equal(kd.cpu().reshape(T, HK, D), rk), "RoPE(k) must be bit-exact"
    # v untouched
    assert torch.equal(qkv_d[:, (HQ + HK) * D:].cpu(), qkv[:, (HQ + HK) * D:])


def test_rope_cos_sin_cache_values():
    # cache layout + dtype the CUDA kernel receives (rotary_embedding.cpp:47-52): [cos_half | sin_half] in bf16
    cs = O.compute_cos_sin_cache(128, 16, 1000000, BF16)
    assert cs.shape == (16, 128) and cs.dtype == BF16
    assert torch.all(cs[0, :64] == 1) and torch.all(cs[0, 64:] == 0)


@pytest.mark.parametrize("T,HK,D,BSZ,NB", [(1, 4, 128, 128, 40), (50, 2, 64, 16, 30), (9, 8, 128, 4, 64), (7, 1, 8, 1, 16)])
def test_reshape_paged_cache_bit_exact(T, HK, D, BSZ, NB, built_lib):
    from xllm_b200 import ops
    g = _g(6)
    kv = _randn((T, 3 * HK * D), g)
    k, v = kv[:, HK * D:2 * HK * D].reshape(T, HK, D), kv[:, 2 * HK * D:].reshape(T, HK, D)
    slots = torch.randperm(NB * BSZ, generator=g)[:T].to(torch.int32)
    slots[T // 2] = -1                                                # skipped slot
    kc, vc = _randn((NB, BSZ, HK, D), g), _randn((NB, BSZ, HK, D), g)
    kc_d, vc_d = kc.to(DEV), vc.to(DEV)
    O.reshape_paged_cache(slots, k, v, kc, vc)
    kvd = kv.to(DEV)
    ops.reshape_paged_cache(slots.to(DEV), kvd[:, HK * D:2 * HK * D].view(T, HK, D), kvd[:, 2 * HK * D:].view(T, HK, D),
                            kc_d, vc_d)
    assert torch.equal(kc_d.cpu(), kc) and torch.equal(vc_d.cpu(), vc)


@pytest.mark.parametrize("T,HQ,HK,D,BSZ", [(1, 28, 4, 128, 128), (19, 14, 2, 64, 16)])
def test_rope_and_cache_fused_equals_sequence(T, HQ, HK, D, BSZ, built_lib):
    from xllm_b200 import ops
    g = _g(7)
    NB = 24
    qkv = _randn((T, (HQ + 2 * HK) * D), g)
    pos = torch.randint(0, 2048, (T,), generator=g)
    cs = O.compute_cos_sin_cache(D, 2048, 1000000, BF16)
    slots = torch.randperm(NB * BSZ, generator=g)[:T].to(torch.int32)
    kc, vc = _randn((NB, BSZ, HK, D), g), _randn((NB, BSZ, HK, D), g)
    q = qkv[:, :HQ * D].reshape(T, HQ, D)
    k = qkv[:, HQ * D:(HQ + HK) * D].reshape(T, HK, D)
    v = qkv[:, (HQ + HK) * D:].reshape(T, HK, D)
    rq, rk = O.rotary_embedding(pos, q, k, cs, True)
    kc_d, vc_d = kc.to(DEV), vc.to(DEV)
    O.reshape_paged_cache(slots, rk, v, kc, vc)
    d = qkv.to(DEV)
    ops.rope_and_cache(pos.to(DEV), d[:, :HQ * D], d[:, HQ * D:(HQ + HK) * D], d[:, (HQ + HK) * D:], cs.to(DEV),
                       slots.to(DEV), kc_d, vc_d, True)
    assert torch.equal(d[:, :HQ * D].cpu().reshape(T, HQ, D), rq)
    assert torch.equal(d[:, HQ * D:(HQ + HK) * D].cpu().reshape(T, HK, D), rk)
    assert torch.equal(kc_d.cpu(), kc) and torch.equal(vc_d.cpu(), vc)


@pytest.mark.parametrize("mode", ["silu", "gelu", "gelu_tanh"])
@pytest.mark.parametrize("T,d", [(1, 18944), (40, 4864), (700, 1024), (3, 100)])
def test_act_and_mul(T, d, mode, built_lib):
    from xllm_b200 import ops
    g = _g(8)
    x = _randn((T, 2 * d), g, 2.0)
    ref = O.act_and_mul(x, mode)
    out = torch.empty(T, d, dtype=BF16, device=DEV)
    ops.act_and_mul(out, x.to(DEV), mode)
    # two roundings (bf16(act) then the bf16 product): a 1-ulp difference of the device expf/erff/tanhf vs torch's
    # in the first can become 2 ulps after the second.  The reference's own test uses allclose(5e-3).
    # gelu: 1 + erf(x/sqrt2) cancels for x << 0, so outputs of magnitude ~1e-6 carry the ABSOLUTE error of erff
    # (~6e-8), i.e. many bf16 ulps of a value that small: an absolute floor of 1e-5 (inputs are O(1)) covers it.
    assert_close_bf16(out, ref, ulps=2, what=f"act_and_mul {mode}", atol=1e-5 if mode != "silu" else 0.0)
    frac = (out.cpu() != ref).float().mean().item()
    assert frac < (5e-3 if mode == "silu" else 2e-2), f"{frac:.4f} of elements differ (device vs host transcendental ulp)"


def test_act_and_mul_bad_mode(built_lib):
    from xllm_b200 import ops
    from xllm_b200._lib import XllmB200Error
    with pytest.raises(XllmB200Error):
        ops.act_and_mul(torch.empty(1, 8, dtype=BF16, device=DEV), torch.empty(1, 16, dtype=BF16, device=DEV), "relu")


@pytest.mark.parametrize("interleaved", [False, True])
@pytest.mark.parametrize("T,HQ,HK,D", [(1, 32, 8, 128), (23, 16, 8, 64), (4, 4, 2, 256)])
def test_fused_qk_norm_rope(T, HQ, HK, D, interleaved, built_lib):
    """reference test: fused_qknorm_rope_test.cpp (NeoX 2e-3 / interleaved 2e-2 vs torch); here vs the oracle."""
    from xllm_b200 import ops
    g = _g(2026)
    qkv = _randn((T, (HQ + 2 * HK) * D), g)
    qw, kw = (1 + 0.1 * torch.randn(D, generator=g)).to(BF16), (1 + 0.1 * torch.randn(D, generator=g)).to(BF16)
    pos = torch.randint(0, 1024, (T,), generator=g)
    cs = O.compute_cos_sin_cache(D, 1024, 1000000, BF16)
    ref = O.fused_qk_norm_rope(qkv, HQ, HK, HK, D, 1e-6, qw, kw, cs, interleaved, pos)
    d = qkv.to(DEV)
    ops.fused_qk_norm_rope(d, HQ, HK, HK, D, 1e-6, qw.to(DEV), kw.to(DEV), cs.to(DEV), interleaved, pos.to(DEV))
    assert_close_bf16(d, ref, ulps=1, what="fused_qk_norm_rope")
    assert torch.equal(d[:, (HQ + HK) * D:].cpu(), qkv[:, (HQ + HK) * D:])   # v untouched
