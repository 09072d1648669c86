"""Op-level oracle: torch-CPU restatements of the reference CUDA kernels.

All functions take/return torch CPU tensors; bf16 tensors carry bf16 dtype and
every intermediate rounding of the reference is reproduced explicitly with
`_r()` (= static_cast<scalar_t>(float) for scalar_t = c10::BFloat16, RNE).
TEST INFRASTRUCTURE - see oracle/__init__.py.
"""
import math

import torch

BF16 = torch.bfloat16
F32 = torch.float32
E4M3 = torch.float8_e4m3fn
LOG2E = 1.4426950408889634


def _r(x: torch.Tensor) -> torch.Tensor:
    """round an fp32 tensor through bf16 (RNE) and back to fp32."""
    return x.to(BF16).to(F32)


# --------------------------------------------------------------------------
# K7 RMSNorm  -- xllm/core/kernels/cuda/norm.cu
# --------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """rms_norm_kernel, norm.cu:43-78: fp32 variance, `(scalar_t)(x*rstd) * weight` (bf16 product)."""
    xf = x.to(F32)
    var = (xf * xf).sum(-1, keepdim=True) / x.shape[-1]
    rstd = torch.rsqrt(var + eps)
    return _r(_r(xf * rstd) * weight.to(F32)).to(BF16)


def fused_add_rms_norm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float):
    """fused_add_rms_norm_kernel, norm.cu:80-136 (width 8) / :139-173 (generic): residual <- bf16(x + residual);
    x <- norm(residual).  Returns (normed, new_residual)."""
    z = _r(x.to(F32) + residual.to(F32))
    var = (z * z).sum(-1, keepdim=True) / x.shape[-1]
    rstd = torch.rsqrt(var + eps)
    out = _r(_r(z * rstd) * weight.to(F32))
    return out.to(BF16), z.to(BF16)


def scaled_fp8_conversion(val: torch.Tensor, inv_scale: torch.Tensor) -> torch.Tensor:
    """scaled_fp8_conversion<true>, fp8_quant_utils.cuh:112-129: x*inv_scale, clamp +-448, RNE e4m3."""
    x = val.to(F32) * inv_scale
    return torch.clamp(x, -448.0, 448.0).to(E4M3)


def rms_norm_static_fp8_quant(x, weight, scale, eps):
    """rms_norm_static_fp8_quant_kernel, norm.cu:228-270: the product with weight stays fp32."""
    xf = x.to(F32)
    var = (xf * xf).sum(-1, keepdim=True) / x.shape[-1]
    rstd = torch.rsqrt(var + eps)
    y = _r(xf * rstd) * weight.to(F32)
    return scaled_fp8_conversion(y, 1.0 / scale.to(F32))


def fused_add_rms_norm_static_fp8_quant(x, residual, weight, scale, eps, vectorized=True):
    """norm.cu:283-345 (width-8 path, taken when hidden%8==0 and 16-byte aligned): the normed value is rounded to
    bf16 before the fp8 conversion; the generic path (:350-396) converts the fp32 product.  Returns (fp8, residual)."""
    z = _r(x.to(F32) + residual.to(F32))
    var = (z * z).sum(-1, keepdim=True) / x.shape[-1]
    rstd = torch.rsqrt(var + eps)
    y = _r(z * rstd) * weight.to(F32)
    if vectorized:
        y = _r(y)
    return scaled_fp8_conversion(y, 1.0 / scale.to(F32)), z.to(BF16)


# --------------------------------------------------------------------------
# K6 FP8 quant -- fp8_quant.cu:78-153, fp8_scaled_quantize.cpp:20-48
# --------------------------------------------------------------------------
def static_scaled_fp8_quant(x, scale):
    return scaled_fp8_conversion(x, 1.0 / scale.to(F32))


def fp8_scaled_quantize(x, scale=None):
    """fp8_scaled_quantize.cpp:36-41.  Dynamic scale: `(amax / 448.0f).clamp_min(1e-12f).to(kFloat32)` where amax is a
    0-dim tensor of the INPUT dtype, so the division and the clamp round to bf16 before the cast to fp32."""
    if scale is None:
        scale = (x.abs().max() / 448.0).clamp_min(1e-12).to(F32).reshape(1)
    return static_scaled_fp8_quant(x, scale), scale


# --------------------------------------------------------------------------
# K9 RoPE -- rope.cu:27-137; cache: rotary_embedding_util.cpp:115-144,303-346, rotary_embedding.cpp:31-52
# --------------------------------------------------------------------------
def compute_inv_freq(rotary_dim: int, rope_theta) -> torch.Tensor:
    """rotary_embedding_util.cpp:303-310; rope_theta is passed as int64 (rotary_embedding.cpp:33)."""
    sl = torch.arange(0, rotary_dim, 2, dtype=F32)
    return 1.0 / torch.pow(torch.tensor(float(int(rope_theta)), dtype=F32), sl / float(rotary_dim))


def llama3_inv_freq(inv_freq: torch.Tensor, factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_pos=8192):
    """Llama-3.1 "llama3" rope_type, restated element by element from the published algorithm (Hugging Face
    modeling_rope_utils._compute_llama3_parameters): not on the reference's path (its llama.h parses the fields,
    models/llm/npu/llama.h:341-346, and ignores them, :77-106) - spec for integration/patches/0002."""
    out = []
    low_wl, high_wl = original_max_pos / low_freq_factor, original_max_pos / high_freq_factor
    for f in inv_freq.tolist():
        wl = 2 * math.pi / f
        if wl < high_wl:
            out.append(f)
        elif wl > low_wl:
            out.append(f / factor)
        else:
            smooth = (original_max_pos / wl - low_freq_factor) / (high_freq_factor - low_freq_factor)
            out.append((1 - smooth) * f / factor + smooth * f)
    return torch.tensor(out, dtype=F32)


def compute_cos_sin_cache(rotary_dim: int, max_pos: int, rope_theta, dtype=BF16, interleaved=False) -> torch.Tensor:
    """[max_pos, rotary_dim] = [cos(rot/2) | sin(rot/2)]: the pre-sliced view the CUDA kernel receives
    (rotary_embedding.cpp:47-52 takes chunks 0 and 2 of cat(cos(cat(f,f)), sin(cat(f,f))))."""
    inv_freq = compute_inv_freq(rotary_dim, rope_theta)
    t = torch.arange(max_pos, dtype=F32)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    return torch.cat([freqs.cos(), freqs.sin()], dim=-1).to(dtype)


def rotary_embedding(positions, q, k, cos_sin_cache, is_neox=True):
    """rotary_embedding_kernel, rope.cu:27-137.  q [T, Hq, D], k [T, Hk, D] (or None).  Every multiply and the
    final add/sub round to bf16 (`x * cos - y * sin` on c10::BFloat16).  Returns new (q, k)."""
    rot = cos_sin_cache.shape[-1]
    emb = rot // 2
    cs = cos_sin_cache[positions.long()].to(F32)        # [T, rot]
    cos, sin = cs[:, None, :emb], cs[:, None, emb:]

    def apply(a):
        if a is None:
            return None
        a = a.clone()
        af = a.to(F32)
        if is_neox:
            x, y = af[..., :emb], af[..., emb:rot]
        else:
            x, y = af[..., 0:rot:2], af[..., 1:rot:2]
        nx = _r(_r(x * cos) - _r(y * sin))
        ny = _r(_r(y * cos) + _r(x * sin))
        if is_neox:
            a[..., :emb] = nx.to(a.dtype)
            a[..., emb:rot] = ny.to(a.dtype)
        else:
            a[..., 0:rot:2] = nx.to(a.dtype)
            a[..., 1:rot:2] = ny.to(a.dtype)
        return a

    return apply(q), apply(k)


def fused_qk_norm_rope(qkv, hq, hk, hv, head_dim, eps, q_weight, k_weight, cos_sin_cache, interleaved, positions):
    """fused_qknorm_rope_kernel, fused_qknorm_rope.cu:84-300: fp32 per-head RMSNorm (e *= rstd*w), fp32 RoPE,
    one rounding at the store.  qkv [T, (hq+hk+hv)*D]; returns a new tensor."""
    T = qkv.shape[0]
    out = qkv.clone().view(T, hq + hk + hv, head_dim)
    x = out[:, :hq + hk].to(F32)
    w = torch.cat([q_weight.to(F32).expand(hq, head_dim), k_weight.to(F32).expand(hk, head_dim)], 0)
    rstd = torch.rsqrt((x * x).sum(-1, keepdim=True) / head_dim + eps)
    e = x * (rstd * w[None])
    rot = cos_sin_cache.shape[-1]
    emb = rot // 2
    cs = cos_sin_cache[positions.long()].to(F32)
    cos, sin = cs[:, None, :emb], cs[:, None, emb:]
    if interleaved:
        a, b = e[..., 0:rot:2].clone(), e[..., 1:rot:2].clone()
        e[..., 0:rot:2] = a * cos - b * sin
        e[..., 1:rot:2] = a * sin + b * cos
    else:
        a, b = e[..., :emb].clone(), e[..., emb:rot].clone()
        e[..., :emb] = a * cos + (-b) * sin
        e[..., emb:rot] = b * cos + a * sin
    out[:, :hq + hk] = e.to(qkv.dtype)
    return out.view(T, -1)


# --------------------------------------------------------------------------
# K11 act_and_mul -- activation.cu:45-130
# --------------------------------------------------------------------------
def act_and_mul(x: torch.Tensor, act_mode: str = "silu") -> torch.Tensor:
    d = x.shape[-1] // 2
    g, u = x[..., :d].to(F32), x[..., d:].to(F32)
    if act_mode == "silu":
        a = g / (1.0 + torch.exp(-g))                                  # :97-102
    elif act_mode == "gelu":
        a = g * 0.5 * (1.0 + torch.erf(g * 0.7071067811865476))        # :104-112
    elif act_mode in ("gelu_tanh", "gelu_pytorch_tanh"):
        inner = 0.7978845608028654 * (g + 0.044715 * g * g * g)        # :114-125
        a = 0.5 * g * (1.0 + torch.tanh(inner))
    else:
        raise ValueError(f"Unsupported act mode: {act_mode}")
    return _r(_r(a) * u).to(BF16)


# --------------------------------------------------------------------------
# K12 KV scatter -- reshape_paged_cache.cu:23-62 (CPU reference: tests/core/kernels/dcu/reshape_paged_cache_test.cpp:33-50)
# --------------------------------------------------------------------------
def reshape_paged_cache(slot_ids, keys, values, key_cache, value_cache):
    """in place on the caches [n_blocks, block_size, Hkv, D]; slot<0 skipped."""
    bs = key_cache.shape[1]
    for t, slot in enumerate(slot_ids.tolist()):
        if slot < 0:
            continue
        key_cache[slot // bs, slot % bs] = keys[t]
        value_cache[slot // bs, slot % bs] = values[t]


# --------------------------------------------------------------------------
# K1/K2/K3 attention.  Math ladder of FlashInfer v0.6.2 fa2 kernels the reference dispatches to
# (batch_decode.cpp:64-84, batch_chunked_prefill.cpp:63-91, batch_prefill.cpp:100-128): fp32 scores, base-2
# softmax with sm_scale*log2(e), P rounded to the q dtype, denominator = sum of the ROUNDED P
# (flashinfer/attention/prefill.cuh compute_sfm_v: m16k16_rowsum on s_frag_f16), fp32 PV accumulate, bf16 out.
# Mask: causal with kv offset, kv_idx <= kv_len - qo_len + q_idx (prefill.cuh:1017).
# Structure (GQA by head grouping, fp32 softmax, cast, PV) as run_eager_causal_padded_attention,
# xllm/core/layers/cuda/flashinfer_attention.cpp:33-89.
# --------------------------------------------------------------------------
def _attend(q, k, v, sm_scale, causal, return_lse=False, round_p=True):
    """q [Lq, Hq, D], k/v [Lk, Hkv, D] (bf16) -> [Lq, Hq, D] bf16 (+ base-2 lse [Lq, Hq]).
    round_p=False keeps the softmax numerators in fp32 (NOT the reference ladder: the "exact math" yardstick the tests use
    to measure how far ONE bf16-P implementation sits from the un-rounded result)."""
    Lq, Hq, D = q.shape
    Lk, Hkv, _ = k.shape
    g = Hq // Hkv
    qf = q.to(F32).view(Lq, Hkv, g, D).permute(1, 2, 0, 3).reshape(Hkv, g * Lq, D)
    kf = k.to(F32).permute(1, 0, 2)                      # [Hkv, Lk, D]
    vf = v.to(F32).permute(1, 0, 2)
    if Lk == 0:
        out = torch.zeros(Lq, Hq, D, dtype=q.dtype)
        return (out, torch.full((Lq, Hq), -math.inf)) if return_lse else out
    s = torch.bmm(qf, kf.transpose(1, 2)) * (sm_scale * LOG2E)   # [Hkv, g*Lq, Lk]
    if causal:
        qi = torch.arange(Lq).repeat(g)                 # row -> q index
        allowed = torch.arange(Lk)[None, :] <= (Lk - Lq + qi)[:, None]
        s = s.masked_fill(~allowed[None], -math.inf)
    m = s.max(-1, keepdim=True).values
    m_safe = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    p = torch.exp2(s - m_safe)
    if round_p:
        p = _r(p)
    l = p.sum(-1, keepdim=True)
    o = torch.bmm(p, vf) / torch.where(l > 0, l, torch.ones_like(l))
    o = o.view(Hkv, g, Lq, D).permute(2, 0, 1, 3).reshape(Lq, Hq, D).to(q.dtype)
    if return_lse:
        lse = (m + torch.log2(l)).view(Hkv, g, Lq).permute(2, 0, 1).reshape(Lq, Hq)
        return o, lse
    return o


def gather_paged_kv(cache, kv_indptr, kv_indices, kv_last_page_len, b):
    """pages of request b -> [kv_len, Hkv, D] (cache [n_blocks, block_size, Hkv, D], NHD: kv_cache_shape.cpp:259-267)."""
    p0, p1 = int(kv_indptr[b]), int(kv_indptr[b + 1])
    if p1 == p0:
        return cache[:0].reshape(0, cache.shape[2], cache.shape[3])
    pages = kv_indices[p0:p1].long()
    bs = cache.shape[1]
    kv_len = (p1 - p0 - 1) * bs + int(kv_last_page_len[b])
    return cache[pages].reshape(-1, cache.shape[2], cache.shape[3])[:kv_len]


def paged_attention(q, k_cache, v_cache, qo_indptr, kv_indptr, kv_indices, kv_last_page_len, sm_scale,
                    causal, return_lse=False, round_p=True):
    """batch_decode (qo_indptr = arange, causal False) / batch_chunked_prefill (ragged q, causal True)."""
    out = torch.empty_like(q)
    lse = torch.empty(q.shape[0], q.shape[1], dtype=F32)
    B = len(kv_indptr) - 1
    for b in range(B):
        q0, q1 = int(qo_indptr[b]), int(qo_indptr[b + 1])
        if q1 == q0:
            continue
        k = gather_paged_kv(k_cache, kv_indptr, kv_indices, kv_last_page_len, b)
        v = gather_paged_kv(v_cache, kv_indptr, kv_indices, kv_last_page_len, b)
        o, ls = _attend(q[q0:q1], k, v, sm_scale, causal, return_lse=True, round_p=round_p)
        out[q0:q1] = o
        lse[q0:q1] = ls
    return (out, lse) if return_lse else out


def ragged_prefill_attention(q, k, v, q_cu_seq_lens, kv_cu_seq_lens, sm_scale, causal=True):
    """batch_prefill (K3): contiguous ragged q/k/v, causal self-attention."""
    out = torch.empty_like(q)
    for b in range(len(q_cu_seq_lens) - 1):
        q0, q1 = int(q_cu_seq_lens[b]), int(q_cu_seq_lens[b + 1])
        k0, k1 = int(kv_cu_seq_lens[b]), int(kv_cu_seq_lens[b + 1])
        if q1 > q0:
            out[q0:q1] = _attend(q[q0:q1], k[k0:k1], v[k0:k1], sm_scale, causal)
    return out


# --------------------------------------------------------------------------
# linears -- layers/common/linear.cpp:616-716,1084-1146,1405-1522 -> kernels/cuda/matmul.cpp:20-24 (F::linear)
# --------------------------------------------------------------------------
def linear(x, w, bias=None):
    """bf16 x bf16 -> fp32 accumulate -> (+bias fp32) -> bf16 (cuBLASLt epilogue order)."""
    y = x.to(F32) @ w.to(F32).t()
    if bias is not None:
        y = y + bias.to(F32)
    return y.to(BF16)


def fp8_scaled_matmul(a8, b8, a_scale, b_scale, bias=None, out_dtype=BF16):
    """cutlass_scaled_mm (cutlass_w8a8/scaled_mm_entry.cu:55-108; ScaledEpilogue[Bias] in
    cutlass_extensions/epilogue/scaled_mm_epilogues_c3x.hpp): D = a_scale * (b_scale * acc) (+ bias), fp32,
    then cast.  a8 [M,K], b8 [N,K] e4m3; scales numel 1 (per-tensor) or [M,1] / [N] (per-token / per-channel)."""
    acc = a8.to(F32) @ b8.to(F32).t()
    bs = b_scale.to(F32).reshape(1, -1)
    as_ = a_scale.to(F32).reshape(-1, 1)
    y = as_ * (bs * acc)
    if bias is not None:
        y = y + bias.to(F32)
    return y.to(out_dtype)


def fp8_linear(x, w8, w_scale, input_scale=None, bias=None):
    """fp8_linear_forward, linear.cpp:137-182: quantise x (static scale if given, else dynamic per-tensor) unless it
    already is e4m3, then scaled matmul."""
    if x.dtype == E4M3:
        a8, a_scale = x, input_scale
    else:
        a8, a_scale = fp8_scaled_quantize(x, input_scale)
    return fp8_scaled_matmul(a8, w8, a_scale, w_scale, bias)
