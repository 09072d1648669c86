"""The reference's own CUDA kernel tests (tests/core/kernels/cuda/*.cpp), restated on CPU with the ORACLE in the place of the
kernel: whatever property the reference asserts of its kernel must hold of the restatement that the GPU parity tests compare
against (pin the oracle to every known-answer test the reference holds for the path).  Sizes, seeds (2026), input
scaling, torch reference expressions and tolerances follow the cited test; nothing here touches the CUDA library."""
import pytest
import torch

from oracle import ops as O

F32, F16, BF16, E4M3 = torch.float32, torch.float16, torch.bfloat16, torch.float8_e4m3fn


def _gen(seed=2026):
    return torch.Generator().manual_seed(seed)


# ---- fp8_quant_test.cpp:44-376 (StaticScaledFP8QuantTest) ------------------------------------------------------------------
@pytest.mark.parametrize("dtype,T,H,scale", [(F32, 128, 256, 1.0), (F16, 64, 512, 0.5), (BF16, 32, 1024, 2.0)])
def test_static_fp8_quant_basic_dtypes(dtype, T, H, scale):
    x = torch.randn(T, H, generator=_gen()).to(dtype)                     # :44-152 Basic / Float16 / BFloat16 input tests
    out = O.static_scaled_fp8_quant(x, torch.tensor([scale]))
    assert out.dtype == E4M3 and out.shape == (T, H)
    f = out.float()
    assert torch.isfinite(f).all()
    assert (x.float().sign() == f.sign()).float().mean().item() > 0.9    # :78-83 sign match ratio


def test_static_fp8_quant_scales_sizes_and_3d():
    x = torch.randn(64, 256, generator=_gen()) * 0.1                      # :154-187 DifferentScalesTest
    for s in (0.1, 0.5, 1.0, 2.0, 10.0):
        assert torch.isfinite(O.static_scaled_fp8_quant(x, torch.tensor([s])).float()).all()
    for T, H in ((1, 64), (16, 128), (64, 256), (128, 512), (256, 1024), (512, 4096)):     # :189-219 DifferentSizesTest
        assert O.static_scaled_fp8_quant(torch.randn(T, H, generator=_gen()), torch.tensor([1.0])).shape == (T, H)
    x3 = torch.randn(4, 32, 128, generator=_gen())                        # :221-252 BatchedTensor3DTest
    o3 = O.static_scaled_fp8_quant(x3, torch.tensor([1.0]))
    assert o3.shape == (4, 32, 128) and torch.isfinite(o3.float()).all()
    assert torch.equal(o3.view(128, 128).view(torch.uint8), O.static_scaled_fp8_quant(x3.view(128, 128), torch.tensor([1.0])).view(torch.uint8))


def test_static_fp8_quant_accuracy_saturation_zeros():
    x = torch.randn(64, 128, generator=_gen()) * 0.5                      # :254-290 QuantizationAccuracyTest
    err = (O.static_scaled_fp8_quant(x, torch.tensor([1.0])).float() - x).abs()
    assert err.mean().item() < 0.5
    assert (err <= x.abs() * 2.0 ** -4 + 2.0 ** -10).all()                # what e4m3 RNE actually guarantees: half an ulp (3-bit mantissa)
    big = torch.randn(32, 64, generator=_gen()) * 10.0                    # :292-327 LargeInputWithScaleTest
    s = max(big.abs().max().item() / 448.0, 1.0)
    assert O.static_scaled_fp8_quant(big, torch.tensor([s])).float().abs().max().item() <= 450.0
    sat = O.static_scaled_fp8_quant(torch.tensor([[3000.0, -3000.0, 448.0, 464.0]]), torch.tensor([1.0])).float()
    assert sat.tolist() == [[448.0, -448.0, 448.0, 448.0]]                # fp8_quant_utils.cuh:112-129: clamp before the conversion
    z = O.static_scaled_fp8_quant(torch.zeros(16, 32), torch.tensor([1.0]))   # :329-350 ZeroValuesTest
    assert (z.float() == 0).all()


# ---- activation_test.cpp:30-140 (ActAndMulKernelTest.MatchesTorchReference) --------------------------------------------------
@pytest.mark.parametrize("dtype", [BF16])     # the test's fp16 leg is out of scope: library and oracle implement the serving dtype only
@pytest.mark.parametrize("mode", ["silu", "gelu", "gelu_tanh"])
@pytest.mark.parametrize("d", [3, 64, 129])
def test_act_and_mul_matches_the_reference_tests_torch_expression(dtype, mode, d):
    x = (torch.randn(4, 7, 2 * d, generator=_gen()) * 0.5).to(dtype)
    a, b = x[..., :d], x[..., d:]
    if mode == "silu":
        ref = (a * torch.sigmoid(a)) * b                                  # evaluated in the tensor dtype, as the test does
    else:
        ref = torch.nn.functional.gelu(a, approximate="none" if mode == "gelu" else "tanh") * b
    out = O.act_and_mul(x, mode)
    assert out.dtype == dtype and out.shape == (4, 7, d)
    assert torch.allclose(out.float(), ref.float(), rtol=5e-3, atol=5e-3)


# ---- cutlass_scaled_mm_test.cpp:44-250 ----------------------------------------------------------------------------------------
def _fp8_pair(M, N, K, g, a_s=None, b_s=None):
    a, b = torch.randn(M, K, generator=g) * 0.5, torch.randn(K, N, generator=g) * 0.5
    a8 = (a if a_s is None else a / a_s[:, None]).to(E4M3)
    b8 = (b if b_s is None else b / b_s[None, :]).to(E4M3)
    return a, b, a8, b8.t().contiguous()                                  # the oracle takes the [N, K] weight (= the column-major b)


def test_scaled_mm_basic_bias_scaling_sizes():
    g = _gen()
    one = torch.ones(1)
    a, b, a8, w8 = _fp8_pair(128, 256, 512, g)                            # :44-95 BasicFP8W8A8Test
    c = O.fp8_scaled_matmul(a8, w8, one, one)
    assert c.shape == (128, 256) and (c.float() - a @ b).abs().max().item() < 2.0
    bias = (torch.randn(256, generator=g) * 0.5).to(BF16)                 # :97-149 FP8W8A8WithBiasTest
    cb = O.fp8_scaled_matmul(a8, w8, one, one, bias)
    assert (cb.float() - (a @ b + bias.float()[None])).abs().max().item() < 2.0
    a_s, b_s = torch.rand(128, generator=g) * 0.1 + 0.9, torch.rand(256, generator=g) * 0.1 + 0.9     # :151-205 WithScalingTest
    a, b, a8, w8 = _fp8_pair(128, 256, 512, g, a_s, b_s)
    d = (O.fp8_scaled_matmul(a8, w8, a_s, b_s).float() - a @ b).abs()
    assert d.max().item() < 2.0 and d.mean().item() < 0.5
    for M, N, K in ((16, 32, 64), (128, 128, 128), (256, 512, 384), (512, 1024, 768)):      # :207-250 DifferentSizesTest
        a, b, a8, w8 = _fp8_pair(M, N, K, g)
        assert O.fp8_scaled_matmul(a8, w8, one, one).shape == (M, N)


# ---- fused_qknorm_rope_test.cpp:25-230 ------------------------------------------------------------------------------------------
def _rope_ref(x, cos, sin, rot, interleaved):
    T, Hh, _ = x.shape
    xf = x.float()
    xr, half = xf[..., :rot], rot // 2
    c, s = cos.view(T, 1, half), sin.view(T, 1, half)
    if interleaved:
        p = xr.reshape(T, Hh, half, 2)
        e, o = p[..., 0], p[..., 1]
        r = torch.stack([e * c - o * s, e * s + o * c], -1).reshape(T, Hh, rot)
    else:
        f, sec = xr[..., :half], xr[..., half:rot]
        r = torch.cat([f * c - sec * s, sec * c + f * s], -1)
    return torch.cat([r, xf[..., rot:]], -1).to(x.dtype)


def _qknorm_rope_ref(qkv, hq, hk, D, eps, qw, kw, cache, interleaved, pos):
    """the torch reference inside the test (:63-107): fp32 norm, ONE rounding to the tensor dtype, then fp32 RoPE, rounded again"""
    T = qkv.shape[0]
    out = qkv.clone()
    q, k = out[:, :hq * D].view(T, hq, D), out[:, hq * D:(hq + hk) * D].view(T, hk, D)
    rot = cache.shape[1]
    cs = cache[pos]
    for t, w in ((q, qw), (k, kw)):
        f = t.float()
        n = (f * torch.rsqrt((f * f).mean(-1, keepdim=True) + eps) * w.float().view(1, 1, D)).to(t.dtype)
        t.copy_(_rope_ref(n, cs[:, :rot // 2], cs[:, rot // 2:], rot, interleaved))
    return out


@pytest.mark.parametrize("T,hq,hk,D,maxpos,dtype,scale,interleaved,tol", [(17, 8, 4, 128, 512, F16, 0.2, False, 2e-3),
                                                                         (11, 6, 2, 64, 256, BF16, 0.15, True, 2e-2)])
def test_fused_qk_norm_rope_matches_the_reference_tests_torch_implementation(T, hq, hk, D, maxpos, dtype, scale, interleaved, tol):
    g = _gen()
    qkv = (torch.randn(T, (hq + 2 * hk) * D, generator=g) * scale).to(dtype)
    qw, kw = torch.randn(D, generator=g).to(dtype), torch.randn(D, generator=g).to(dtype)
    cache = torch.randn(maxpos, D, generator=g)                           # the test fills the cache with N(0,1) floats
    pos = torch.randint(0, maxpos, (T,), generator=g)
    ref = _qknorm_rope_ref(qkv, hq, hk, D, 1e-6, qw, kw, cache, interleaved, pos)
    out = O.fused_qk_norm_rope(qkv, hq, hk, hk, D, 1e-6, qw, kw, cache, interleaved, pos)
    assert torch.equal(out[:, (hq + hk) * D:], qkv[:, (hq + hk) * D:]), "v must pass through untouched"
    assert torch.allclose(out.float(), ref.float(), rtol=tol, atol=tol)


# ---- tests/core/kernels/dcu/*.cpp: the mirrors of the same norm.cu / rope.cu / reshape_paged_cache.cu / matmul sources, with CPU
# ---- torch references inside the tests.  bf16 legs only (the serving dtype); the 12 x 8192 x 4096 cases run as 2 x 8 x 4096 ------
@pytest.mark.parametrize("name,shape,zero,strided", [("SmallHidden", (64, 64), False, False), ("LargeHidden", (64, 4096), False, False),
                                                      ("SingleToken", (1, 256), False, False), ("ZeroInput", (64, 256), True, False),
                                                      ("StridedInput", (64, 256), False, True), ("3D", (2, 8, 4096), False, False)])
def test_rms_norm_dcu_cases(name, shape, zero, strided):
    g = _gen()                                                            # norm_test.cpp:82-148
    H = shape[-1]
    if zero:
        x = torch.zeros(shape, dtype=BF16)
    elif strided:
        x = torch.randn(shape[0], 2 * H, generator=g).to(BF16)[:, :H]     # stride(-2) != hidden
    else:
        x = torch.randn(shape, generator=g).to(BF16)
    w = (torch.randn(H, generator=g) * 0.5 + 1.0).to(BF16)
    out = O.rms_norm(x, w, 1e-6)
    ref = torch.nn.functional.rms_norm(x, (H,), w, 1e-6)
    assert out.shape == x.shape and torch.isfinite(out.float()).all()
    assert torch.allclose(out.float(), ref.float(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("shape", [(64, 256), (64, 100), (64, 4096), (8192, 128), (2, 8, 4096)])
def test_fused_add_rms_norm_dcu_cases(shape):
    g = _gen()                                                            # norm_test.cpp:190-262
    H = shape[-1]
    x = (torch.randn(shape, generator=g) * 0.3).to(BF16)
    r = (torch.randn(shape, generator=g) * 0.3).to(BF16)
    w = torch.randn(H, generator=g).to(BF16)
    out, res = O.fused_add_rms_norm(x, r, w, 1e-6)
    upd = (x.float() + r.float()).to(BF16)
    ref = torch.nn.functional.rms_norm(upd, (H,), w, 1e-6)
    assert torch.equal(res, upd), "residual += input is one fp32 add rounded once"
    assert torch.allclose(out.float(), ref.float(), rtol=1e-2, atol=1e-2)


def _rope_cpu_reference(pos, x, cache, is_neox):
    """apply_rope_reference_cpu (rope_test.cpp:72-150): fp32, x*c - y*s / y*c + x*s on the pairs (off, emb+off) or (2off, 2off+1)"""
    T, Hh, D = x.shape
    emb = cache.shape[1] // 2
    cs = cache[pos]
    c, s = cs[:, None, :emb], cs[:, None, emb:]
    out = x.clone()
    if is_neox:
        a, b = x[..., :emb], x[..., emb:2 * emb]
        out[..., :emb], out[..., emb:2 * emb] = a * c - b * s, b * c + a * s
    else:
        a, b = x[..., 0::2], x[..., 1::2]
        out[..., 0::2], out[..., 1::2] = a * c - b * s, b * c + a * s
    return out


def _rope_bound(x, is_neox, emb):
    """all-bf16 arithmetic (rope.cu:27-54 on c10::BFloat16: the cache value, the products and the sum each round once, unit roundoff
    u = 2^-8): every output is within 3 u (|x| + |y|) of the exact rotation of its pair (x, y)"""
    ax = x.abs()
    if is_neox:
        pair = ax[..., :emb] + ax[..., emb:2 * emb]
        return torch.cat([pair, pair], -1) * 3 * 2.0 ** -8
    pair = ax[..., 0::2] + ax[..., 1::2]
    return torch.repeat_interleave(pair, 2, dim=-1) * 3 * 2.0 ** -8


# (is_neox, with_key, tokens, heads, kv heads, head size) of rope_test.cpp:395-770; the layouts / 2-D positions of the grid are views of
# the same [T, H, D] problem.  The test's bf16 leg is the (6, 8, 2, 16) case at rtol = atol = 1e-2; its two Smoke cases are fp32 there
@pytest.mark.parametrize("is_neox,with_key,T,hq,hk,D,ref_tol", [(True, True, 4, 2, 2, 8, True), (True, True, 6, 8, 2, 16, True),
                                                                (False, False, 5, 2, 2, 8, True), (False, True, 5, 6, 2, 8, True),
                                                                (True, True, 3, 4, 2, 8, True), (True, False, 5, 4, 4, 8, True),
                                                                (False, True, 6, 6, 2, 8, True), (True, True, 1, 1, 1, 2, True),
                                                                (False, False, 1, 2, 2, 8, True), (True, True, 32, 16, 4, 64, False),
                                                                (True, True, 128, 32, 8, 64, False)])
def test_rotary_embedding_dcu_cases(is_neox, with_key, T, hq, hk, D, ref_tol):
    g = _gen()                                                            # rope_test.cpp:286-392 run_case, cases :395-770
    max_pos = max(32, T + 8)
    pos = torch.tensor([(i * 3 + 1) % max_pos for i in range(T)])          # make_positions_1d_cpu
    i = torch.arange(D // 2, dtype=torch.float32)
    theta = (torch.arange(max_pos, dtype=torch.float32)[:, None] + 1) * (i[None] + 1) * 0.01      # make_cos_sin_cache_cpu :40-56
    cache = torch.cat([torch.cos(theta), torch.sin(theta)], 1)
    q = torch.randn(T, hq, D, generator=g).to(BF16)
    k = torch.randn(T, hk, D, generator=g).to(BF16) if with_key else None
    q2, k2 = O.rotary_embedding(pos, q, k, cache.to(BF16), is_neox)       # the kernel gets the cache in the tensor dtype (:348)
    for got, x in ((q2, q), (k2, k)):
        if x is None:
            assert got is None
            continue
        ref = _rope_cpu_reference(pos, x.float(), cache, is_neox)
        assert ((got.float() - ref).abs() <= _rope_bound(x.float(), is_neox, D // 2)).all()
        if ref_tol:
            assert torch.allclose(got.float(), ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("n_tokens,n_blocks,block_size,hkv,D", [(4, 1, 16, 1, 64), (32, 8, 16, 4, 128), (64, 4, 64, 8, 128),
                                                                (256, 16, 64, 8, 128), (1, 4, 16, 4, 128)])
def test_reshape_paged_cache_dcu_cases(n_tokens, n_blocks, block_size, hkv, D):
    g = _gen()                                                            # reshape_paged_cache_test.cpp:33-50, :135-155
    keys = torch.randn(n_tokens, hkv, D, generator=g).to(BF16)
    vals = torch.randn(n_tokens, hkv, D, generator=g).to(BF16)
    slots = torch.randperm(n_blocks * block_size, generator=g)[:n_tokens].to(torch.int32)
    kc = torch.zeros(n_blocks, block_size, hkv, D, dtype=BF16)
    vc = torch.zeros_like(kc)
    O.reshape_paged_cache(slots, keys, vals, kc, vc)
    rk, rv = torch.zeros_like(kc), torch.zeros_like(vc)
    for t in range(n_tokens):
        b, o = int(slots[t]) // block_size, int(slots[t]) % block_size
        rk[b, o], rv[b, o] = keys[t], vals[t]
    assert torch.equal(kc, rk) and torch.equal(vc, rv)


@pytest.mark.parametrize("shape_a,shape_b,with_bias", [((1, 1024), (1024, 1024), False), ((128, 512), (256, 512), True),
                                                       ((2, 64, 256), (512, 256), False), ((33, 96), (40, 96), True)])
def test_matmul_dcu_cases(shape_a, shape_b, with_bias):
    g = _gen()                                                            # matmul_test.cpp:124-150: F::linear on CPU, bf16 tol 2e-2
    a = (torch.randn(shape_a, generator=g) * 0.3).to(BF16)
    b = (torch.randn(shape_b, generator=g) * 0.3).to(BF16)
    bias = (torch.randn(shape_b[0], generator=g) * 0.1).to(BF16) if with_bias else None
    out = O.linear(a.reshape(-1, shape_a[-1]), b, bias).reshape(*shape_a[:-1], shape_b[0])
    ref = torch.nn.functional.linear(a, b, bias)
    assert torch.allclose(out.float(), ref.float(), rtol=2e-2, atol=2e-2)


def test_e4m3_conversion_is_rne_saturating_bit_level():
    """scaled_fp8_conversion (fp8_quant_utils.cuh:78-129) = clamp to +-448 then __nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3):
    round-to-nearest-even onto the e4m3fn grid (3 mantissa bits, exponent bias 7, subnormal step 2^-9, max 448).  The oracle leans
    on torch's float8_e4m3fn cast for the rounding; this checks that cast against an independent construction of the grid: every
    representable value, every midpoint between neighbours (ties to the even code), and values just either side of the midpoints."""
    import numpy as np
    codes = np.arange(0, 127, dtype=np.uint8)                              # 0x00 .. 0x7e: +0 .. 448 (0x7f is NaN)
    e, m = codes >> 3, (codes & 7).astype(np.float64)
    grid = np.where(e == 0, m * 2.0 ** -9, (1 + m / 8) * 2.0 ** (e.astype(np.float64) - 7))
    assert grid[-1] == 448.0 and grid[1] == 2.0 ** -9 and (np.diff(grid) > 0).all()
    assert torch.equal(torch.tensor(grid, dtype=torch.float32).to(E4M3).view(torch.uint8), torch.tensor(codes))
    mid = (grid[:-1] + grid[1:]) / 2
    want_tie = np.where(codes[:-1] % 2 == 0, codes[:-1], codes[1:])        # ties go to the even code
    got = torch.tensor(mid, dtype=torch.float32).to(E4M3).view(torch.uint8).numpy()
    assert (got == want_tie).all()
    below = torch.tensor(np.nextafter(mid.astype(np.float32), np.float32(0)), dtype=torch.float32).to(E4M3).view(torch.uint8).numpy()
    above = torch.tensor(np.nextafter(mid.astype(np.float32), np.float32(1e9)), dtype=torch.float32).to(E4M3).view(torch.uint8).numpy()
    assert (below == codes[:-1]).all() and (above == codes[1:]).all()
    # through the oracle: scale, clamp, convert - and the sign is carried
    x = torch.tensor([[465.0, -1000.0, 0.3, -0.3, 2.0 ** -10, 2.0 ** -11]])
    out = O.static_scaled_fp8_quant(x, torch.tensor([1.0])).float().tolist()[0]
    assert out == [448.0, -448.0, 0.3125, -0.3125, 0.0, 0.0]             # 2^-10 is the midpoint of 0 and 2^-9: tie to even = 0
