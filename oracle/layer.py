"""Layer/model composition oracle (CUDA build of the reference).

Follows Qwen2AttentionImpl::forward (xllm/core/layers/common/qwen2_attention.cpp:132-193),
FlashInferAttentionImpl::forward (xllm/core/layers/cuda/flashinfer_attention.cpp:112-157: scatter k/v into the
cache first, then prefill | chunked prefill | decode), DenseMLPImpl::forward (layers/common/dense_mlp.cpp:97-118),
Qwen2DecoderLayerImpl::forward / apply_norm (layers/qwen2_decoder_layer.cpp:64-112) and
LlmModelImplBase::forward (xllm/models/llm/llm_model_base.h:60-131).
TEST INFRASTRUCTURE - see oracle/__init__.py.
"""
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch

from . import ops

BF16 = torch.bfloat16


@dataclass
class AttnMeta:
    """The subset of AttentionMetadata (layers/common/attention_metadata.h:73-187) the math needs."""
    is_prefill: bool
    is_chunked_prefill: bool
    q_cu_seq_lens: torch.Tensor           # int32 [B+1]
    kv_cu_seq_lens: torch.Tensor          # int32 [B+1]
    slot_mapping: torch.Tensor            # int32 [T]
    paged_kv_indptr: Optional[torch.Tensor] = None
    paged_kv_indices: Optional[torch.Tensor] = None
    paged_kv_last_page_len: Optional[torch.Tensor] = None


def attention_forward(meta: AttnMeta, q, k, v, k_cache, v_cache, scale, n_heads, n_kv_heads, head_dim):
    """flashinfer_attention.cpp:112-157.  q [T,Hq*D], k/v [T,Hkv*D]; caches NHD, updated in place."""
    T = q.shape[0]
    q3, k3, v3 = q.view(T, n_heads, head_dim), k.view(T, n_kv_heads, head_dim), v.view(T, n_kv_heads, head_dim)
    ops.reshape_paged_cache(meta.slot_mapping, k3, v3, k_cache, v_cache)            # :128-131
    if meta.is_prefill and not meta.is_chunked_prefill:
        out = ops.ragged_prefill_attention(q3, k3, v3, meta.q_cu_seq_lens, meta.kv_cu_seq_lens, scale, causal=True)
    else:
        causal = meta.is_chunked_prefill
        out = ops.paged_attention(q3, k_cache, v_cache, meta.q_cu_seq_lens, meta.paged_kv_indptr,
                                  meta.paged_kv_indices, meta.paged_kv_last_page_len, scale, causal=causal)
    return out.reshape(T, n_heads * head_dim)


class Qwen2AttentionOracle:
    def __init__(self, qkv_w, qkv_b, o_w, n_heads, n_kv_heads, head_dim, cos_sin_cache,
                 linear: Callable = ops.linear, o_linear: Optional[Callable] = None):
        self.qkv_w, self.qkv_b, self.o_w = qkv_w, qkv_b, o_w
        self.n_heads, self.n_kv_heads, self.head_dim = n_heads, n_kv_heads, head_dim
        self.cos_sin_cache = cos_sin_cache
        self.scale = head_dim ** -0.5                                     # qwen2_attention.cpp:41
        self.linear, self.o_linear = linear, (o_linear or linear)

    def forward(self, positions, hidden, meta: AttnMeta, k_cache, v_cache):
        qkv = self.linear(hidden, self.qkv_w, self.qkv_b)                 # :137
        qs, kvs = self.n_heads * self.head_dim, self.n_kv_heads * self.head_dim
        T = qkv.shape[0]
        q = qkv[:, :qs].reshape(T, self.n_heads, self.head_dim)
        k = qkv[:, qs:qs + kvs].reshape(T, self.n_kv_heads, self.head_dim)
        v = qkv[:, qs + kvs:qs + 2 * kvs]
        q, k = ops.rotary_embedding(positions, q, k, self.cos_sin_cache, is_neox=True)   # :173-176
        out = attention_forward(meta, q.reshape(T, qs), k.reshape(T, kvs), v.contiguous(), k_cache, v_cache,
                                self.scale, self.n_heads, self.n_kv_heads, self.head_dim)
        return self.o_linear(out, self.o_w, None)                         # :189


class Qwen2DecoderLayerOracle:
    """qwen2_decoder_layer.cpp:89-112.  pre_fp8_scale / post_fp8_scale: the static input scales of qkv_proj / gate_up_proj of an
    FP8 checkpoint (get_fp8_input_scale, qwen2_attention.cpp:204-209 / dense_mlp.cpp:137-142): apply_norm (:64-84) then takes
    RMSNormImpl::forward_fp8 (rms_norm.cpp:94-128) and hands e4m3 activations to the linear, which skips its own quantisation
    (linear.cpp:150-157).  None = plain norms."""

    def __init__(self, attn: Qwen2AttentionOracle, input_norm_w, post_norm_w, eps, gate_up, down,
                 linear: Callable = ops.linear, pre_fp8_scale=None, post_fp8_scale=None):
        self.attn, self.in_w, self.post_w, self.eps = attn, input_norm_w, post_norm_w, eps
        self.gate_up, self.down = gate_up, down      # callables x -> y
        self.linear = linear
        self.pre_s, self.post_s = pre_fp8_scale, post_fp8_scale

    def _apply_norm(self, x, residual, w, scale):
        vec = x.shape[-1] % 8 == 0                                        # norm.cu:283-345 width-8 path vs :350-396
        if residual is None:                                              # apply_norm :72-79
            h = ops.rms_norm(x, w, self.eps) if scale is None else ops.rms_norm_static_fp8_quant(x, w, scale, self.eps)
            return h, x
        if scale is None:
            return ops.fused_add_rms_norm(x, residual, w, self.eps)
        return ops.fused_add_rms_norm_static_fp8_quant(x, residual, w, scale, self.eps, vectorized=vec)

    def forward(self, x, residual, positions, meta, k_cache, v_cache):
        h, residual = self._apply_norm(x, residual, self.in_w, self.pre_s)
        h = self.attn.forward(positions, h, meta, k_cache, v_cache)
        h, residual = self._apply_norm(h, residual, self.post_w, self.post_s)
        h = self.down(ops.act_and_mul(self.gate_up(h), "silu"))           # dense_mlp.cpp:97-118
        return h, residual
