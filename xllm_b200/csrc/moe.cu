// Mixture-of-experts decode path: router top-k and the gated expert MLP for decode-sized token counts.
//   xb_moe_fused_topk        replaces xllm::kernel::cuda::moe_fused_topk (kernels/cuda/moe/moe_fused_topk.cu:22-58 ->
//                            moe_topk_softmax_kernels.cuh / moe_topk_sigmoid_kernels.cuh): softmax or sigmoid scores in fp32,
//                            optional correction bias (added for the SELECTION, subtracted again from the returned weight,
//                            moe_topk_sigmoid_kernels.cuh:66-71,131-139), top-k by value with ties to the lower expert index
//                            (the ordering tests/core/kernels/cuda/moe/moe_topk_test.cu:31-55 pins), optional
//                            renormalisation by the sum of the k selected weights taken in selection order.
//   xb_moe_experts_bf16      replaces xllm::kernel::cuda::cutlass_fused_moe (kernels/cuda/moe/fused_moe.cpp:23-124 -> FlashInfer
//                            fused_moe_100) for unquantised bf16 experts - the only kind the reference's CUDA FusedMoE accepts
//                            (layers/cuda/fused_moe.cpp:39-42): fc1 [E, 2I, H] in [up | gate] row order (fused_moe.cpp /
//                            layers/cuda/fused_moe.cpp:124-126 "CUTLASS SwiGLU consumes fc1 as [linear, gate]"), fc2 [E, H, I].
// Decode design (HBM-bound): every (token, selected expert) pair streams its expert's weights once with the register-
// fragment GEMV of linear_small_m.cu; fc1 computes the up rows and the matching gate rows in one CTA and applies
// silu(gate) * up on the fp32 sums (one rounding to bf16), fc2 writes bf16 rows per pair, a combine kernel forms
// sum_k scale_k * y2_k in fp32 in selection order.  Experts outside [expert_begin, expert_end) (expert parallelism:
// fused_moe.cpp ep_rank / ep_size) contribute zero, as in the reference where the EP all-reduce adds the other ranks.
#include "common.cuh"

namespace xb {

// ---------------------------------------------------------------------------------------------------------------------
// router: one warp per token, up to 32 * kPerLane experts
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTopkPerLane = 16;   // experts per lane: E <= 512

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(128)
moe_fused_topk_kernel(float* __restrict__ weights, int32_t* __restrict__ ids, const T* __restrict__ gating, int64_t g_stride,
                      const float* __restrict__ bias, int num_tokens, int E, int k, int renormalize, int sigmoid) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * 4 + warp;
  if (tok >= num_tokens) return;
  const T* row = gating + (int64_t)tok * g_stride;
  float sc[kTopkPerLane];      // selection score of expert lane + 32 * i (-inf beyond E)
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kTopkPerLane; ++i) {
    const int e = lane + 32 * i;
    sc[i] = e < E ? to_f32(row[e]) : -INFINITY;
    mx = fmaxf(mx, sc[i]);
  }
  if (sigmoid) {
#pragma unroll
    for (int i = 0; i < kTopkPerLane; ++i) {
      const int e = lane + 32 * i;
      if (e < E) {
        float v = 1.0f / (1.0f + expf(-sc[i]));
        if (bias) v = v + bias[e];
        sc[i] = v;
      }
    }
  } else {
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kTopkPerLane; ++i) {
      const int e = lane + 32 * i;
      if (e < E) {
        sc[i] = expf(sc[i] - mx);
        sum += sc[i];
      }
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < kTopkPerLane; ++i)
      if (lane + 32 * i < E) sc[i] = sc[i] * inv;
  }
  float row_sum = 0.f;
  float my_w = 0.f;            // lane j keeps the j-th selected weight
  int my_id = 0;
  for (int kk = 0; kk < k; ++kk) {
    // lane-local best (ties: lower expert index = lower i), then the warp's best (ties: lower expert index)
    float bv = -INFINITY;
    int be = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < kTopkPerLane; ++i) {
      const int e = lane + 32 * i;
      if (e < E && (sc[i] > bv || (sc[i] == bv && e < be))) {
        bv = sc[i];
        be = e;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oe = __shfl_xor_sync(0xffffffffu, be, o);
      if (ov > bv || (ov == bv && oe < be)) {
        bv = ov;
        be = oe;
      }
    }
    // remove the winner from the candidates
    if ((be & 31) == lane) {
#pragma unroll
      for (int i = 0; i < kTopkPerLane; ++i)
        if (i == (be >> 5)) sc[i] = -INFINITY;
    }
    float w = bv;
    if (sigmoid && bias) w -= bias[be];
    row_sum += w;
    if (lane == kk) {
      my_w = w;
      my_id = be;
    }
  }
  if (lane < k) {
    if (renormalize) my_w = my_w * (1.0f / row_sum);
    weights[(int64_t)tok * k + lane] = my_w;
    ids[(int64_t)tok * k + lane] = my_id;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// expert-indexed GEMV: y[pair] = W[expert(pair)] . x[row(pair)]   (bf16 weights [E][N][K] row-major, M = 1 per pair)
//   kSwiGLU: W is fc1 [E][2I][K] in [up | gate] order; the CTA owns up rows n0.. and gate rows I + n0.. and writes
//            a[pair][n0 + r] = bf16(silu(gate) * up), both from the fp32 sums.
// lane (g, t) reads rows n0+g and n0+g+8, 16 bytes per k32 tile (the layout of linear_bf16_small_m_kernel); the token sits
// in column 0 of the n8 slot.
// ---------------------------------------------------------------------------------------------------------------------
template <bool kSwiGLU>
__global__ void __launch_bounds__(256)
moe_expert_gemv_kernel(__nv_bfloat16* __restrict__ y, int64_t y_stride, const __nv_bfloat16* __restrict__ x, int64_t x_stride,
                       int x_row_div /* x row = pair / x_row_div */, const __nv_bfloat16* __restrict__ w,
                       const int32_t* __restrict__ expert_ids, int N /* output rows per expert (I for SwiGLU) */, int K,
                       int expert_begin, int expert_end) {
  constexpr int kW = 8;
  __shared__ float red[kW][kSwiGLU ? 2 : 1][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int pair = blockIdx.y;
  const int n0 = blockIdx.x * 16;
  pdl_launch_dependents();
  pdl_wait();
  const int e_glob = expert_ids[pair];
  const bool mine = e_glob >= expert_begin && e_glob < expert_end;
  const int e = e_glob - expert_begin;
  const int rows_per_expert = kSwiGLU ? 2 * N : N;
  const __nv_bfloat16* we = w + (int64_t)(mine ? e : 0) * rows_per_expert * K;
  const __nv_bfloat16* xr = x + (int64_t)(pair / x_row_div) * x_stride;
  const int ktiles = K >> 5;
  float acc[kSwiGLU ? 2 : 1][4];
#pragma unroll
  for (int h = 0; h < (kSwiGLU ? 2 : 1); ++h)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[h][i] = 0.f;
  if (mine) {
    const bool r0 = n0 + g < N, r1 = n0 + g + 8 < N;
    constexpr int kU = 4;
    for (int kt0 = warp; kt0 < ktiles; kt0 += kW * kU) {
      uint4 wa[kSwiGLU ? 2 : 1][kU], wb[kSwiGLU ? 2 : 1][kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int kt = kt0 + u * kW;
#pragma unroll
        for (int h = 0; h < (kSwiGLU ? 2 : 1); ++h) {
          wa[h][u] = make_uint4(0, 0, 0, 0);
          wb[h][u] = make_uint4(0, 0, 0, 0);
          if (kt < ktiles) {
            const __nv_bfloat16* base = we + (int64_t)(h * N + n0) * K + (kt << 5) + 8 * t;
            if (r0) wa[h][u] = ldg_stream(base + (int64_t)g * K);
            if (r1) wb[h][u] = ldg_stream(base + (int64_t)(g + 8) * K);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int kt = kt0 + u * kW;
        if (kt < ktiles) {
          uint4 xv = make_uint4(0, 0, 0, 0);
          if (g == 0) xv = *reinterpret_cast<const uint4*>(xr + (kt << 5) + 8 * t);   // token in column 0 only
#pragma unroll
          for (int h = 0; h < (kSwiGLU ? 2 : 1); ++h) {
            mma_bf16_16816(acc[h], wa[h][u].x, wb[h][u].x, wa[h][u].y, wb[h][u].y, xv.x, xv.y);
            mma_bf16_16816(acc[h], wa[h][u].z, wb[h][u].z, wa[h][u].w, wb[h][u].w, xv.z, xv.w);
          }
        }
      }
    }
  }
  // C fragment: c0 = (row g, col 2t), c2 = (row g+8, col 2t); the token is column 0 -> lanes with t == 0
  if (t == 0) {
#pragma unroll
    for (int h = 0; h < (kSwiGLU ? 2 : 1); ++h) {
      red[warp][h][g] = acc[h][0];
      red[warp][h][g + 8] = acc[h][2];
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int r = threadIdx.x;
    if (n0 + r < N) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int ww = 0; ww < kW; ++ww) {
        s0 += red[ww][0][r];
        if (kSwiGLU) s1 += red[ww][1][r];
      }
      float out = s0;
      if (kSwiGLU) out = (s1 / (1.0f + expf(-s1))) * s0;      // silu(gate) * up on the fp32 sums
      y[(int64_t)pair * y_stride + n0 + r] = __float2bfloat16_rn(mine ? out : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// W4A16 experts (SURVEY cfg[4] "MoE W4A16"; additive like the dense W4A16 linears - the reference's CUDA FusedMoE takes
// unquantised experts only): the same expert-indexed GEMV over the tile-packed int4 layout of linear_small_m.cu, per expert
//   qweight [E][rows/16][K/64][32 lanes][4] u32,  meta [E][K/g][rows] u32 (bf16 scale | bf16(128 + zero) << 16),
// dequantised bit-exactly as w = bf16((q - z) * s) (spec form 1 of oracle/quant.py) right before the mma.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t moe_lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(d) : "r"(a), "r"(mask), "r"(orv));
  return d;
}
__device__ __forceinline__ uint32_t moe_hsub2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

template <bool kSwiGLU>
__global__ void __launch_bounds__(256)
moe_expert_gemv_w4_kernel(__nv_bfloat16* __restrict__ y, int64_t y_stride, const __nv_bfloat16* __restrict__ x, int64_t x_stride,
                          int x_row_div, const uint4* __restrict__ qweight, const uint32_t* __restrict__ meta,
                          const int32_t* __restrict__ expert_ids, int N /* output rows per expert (I for SwiGLU) */, int K,
                          int gshift /* log2(k64 tiles per quantisation group) */, int expert_begin, int expert_end) {
  constexpr int kW = 8, kH = kSwiGLU ? 2 : 1;
  __shared__ float red[kW][kH][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int pair = blockIdx.y;
  const int n0 = blockIdx.x * 16;
  pdl_launch_dependents();
  pdl_wait();
  const int e_glob = expert_ids[pair];
  const bool mine = e_glob >= expert_begin && e_glob < expert_end;
  const int e = mine ? e_glob - expert_begin : 0;
  const int rows = kSwiGLU ? 2 * N : N;
  const int ktiles = K >> 6;
  const uint4* qe = qweight + (int64_t)e * (rows >> 4) * ktiles * 32;
  const uint32_t* me = meta + (int64_t)e * ((ktiles >> gshift) * (int64_t)rows);
  const __nv_bfloat16* xr = x + (int64_t)(pair / x_row_div) * x_stride;
  float acc[kH][4];
#pragma unroll
  for (int h = 0; h < kH; ++h)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[h][i] = 0.f;
  if (mine && n0 < N) {
    constexpr int kU = 4;
    for (int kt0 = warp; kt0 < ktiles; kt0 += kW * kU) {
      uint4 wq[kH][kU];
      uint32_t m0[kH][kU], m1[kH][kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int kt = kt0 + u * kW;
#pragma unroll
        for (int h = 0; h < kH; ++h) {
          wq[h][u] = make_uint4(0, 0, 0, 0);
          m0[h][u] = m1[h][u] = 0;
          if (kt < ktiles) {
            const int row0 = h * N + n0;                       // up rows, then the matching gate rows
            wq[h][u] = ldg_stream(qe + ((int64_t)(row0 >> 4) * ktiles + kt) * 32 + lane);
            const uint32_t* mr = me + (int64_t)(kt >> gshift) * rows + row0;
            m0[h][u] = __ldg(mr + g);
            m1[h][u] = __ldg(mr + g + 8);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int kt = kt0 + u * kW;
        if (kt < ktiles) {
          uint4 xlo = make_uint4(0, 0, 0, 0), xhi = make_uint4(0, 0, 0, 0);
          if (g == 0) {                                        // the token sits in column 0 of the n8 slot
            const uint4* xp = reinterpret_cast<const uint4*>(xr + (kt << 6) + 16 * t);
            xlo = xp[0];
            xhi = xp[1];
          }
#pragma unroll
          for (int h = 0; h < kH; ++h) {
            const uint32_t s0 = __byte_perm(m0[h][u], 0, 0x1010), z0 = __byte_perm(m0[h][u], 0, 0x3232);
            const uint32_t s1 = __byte_perm(m1[h][u], 0, 0x1010), z1 = __byte_perm(m1[h][u], 0, 0x3232);
            const uint32_t* wv = &wq[h][u].x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t w = wv[j];
              const uint32_t a0 = hmul2_bf16x2(moe_hsub2(moe_lop3_and_or(w, 0x000f000fu, 0x43004300u), z0), s0);
              const uint32_t a1 = hmul2_bf16x2(moe_hsub2(moe_lop3_and_or(w >> 4, 0x000f000fu, 0x43004300u), z1), s1);
              const uint32_t a2 = hmul2_bf16x2(moe_hsub2(moe_lop3_and_or(w >> 8, 0x000f000fu, 0x43004300u), z0), s0);
              const uint32_t a3 = hmul2_bf16x2(moe_hsub2(moe_lop3_and_or(w >> 12, 0x000f000fu, 0x43004300u), z1), s1);
              const uint32_t* xv = j < 2 ? &xlo.x : &xhi.x;
              mma_bf16_16816(acc[h], a0, a1, a2, a3, xv[(j & 1) * 2], xv[(j & 1) * 2 + 1]);
            }
          }
        }
      }
    }
  }
  if (t == 0) {
#pragma unroll
    for (int h = 0; h < kH; ++h) {
      red[warp][h][g] = acc[h][0];
      red[warp][h][g + 8] = acc[h][2];
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int r = threadIdx.x;
    if (n0 + r < N) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int ww = 0; ww < kW; ++ww) {
        s0 += red[ww][0][r];
        if (kSwiGLU) s1 += red[ww][1][r];
      }
      float out = s0;
      if (kSwiGLU) out = (s1 / (1.0f + expf(-s1))) * s0;
      y[(int64_t)pair * y_stride + n0 + r] = __float2bfloat16_rn(mine ? out : 0.f);
    }
  }
}

// out[t] = bf16( sum_k scale[t,k] * y2[t*k + kk] ), fp32, selection order
__global__ void __launch_bounds__(256)
moe_combine_kernel(__nv_bfloat16* __restrict__ out, int64_t out_stride, const __nv_bfloat16* __restrict__ y2, const float* __restrict__ scales,
                   int k, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int tok = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < H / 2; c += gridDim.x * blockDim.x) {
    float a0 = 0.f, a1 = 0.f;
    for (int kk = 0; kk < k; ++kk) {
      const float s = scales[(int64_t)tok * k + kk];
      const uint32_t v = reinterpret_cast<const uint32_t*>(y2 + ((int64_t)tok * k + kk) * H)[c];
      a0 = fmaf(s, bf16lo(v), a0);
      a1 = fmaf(s, bf16hi(v), a1);
    }
    reinterpret_cast<uint32_t*>(out + (int64_t)tok * out_stride)[c] = pack_bf16x2(a0, a1);
  }
}

}  // namespace xb

using namespace xb;

extern "C" int xb_moe_fused_topk(float* topk_weights, int32_t* topk_ids, const void* gating_output, int gating_is_bf16,
                                 int64_t gating_stride, const float* correction_bias, int num_tokens, int num_experts,
                                 int topk, int renormalize, int scoring_sigmoid, xb_stream_t stream) {
  if (num_tokens == 0) return 0;
  XB_CHECK(num_experts >= 1 && num_experts <= 32 * kTopkPerLane, "moe_fused_topk: %d experts unsupported (1..%d)", num_experts,
           32 * kTopkPerLane);
  XB_CHECK(topk >= 1 && topk <= 32 && topk <= num_experts, "moe_fused_topk: topk %d out of range (1..min(32, experts))", topk);
  XB_CHECK(scoring_sigmoid || correction_bias == nullptr, "moe_fused_topk: correction_bias goes with sigmoid scoring");
  dim3 grid((num_tokens + 3) / 4), block(128);
  if (gating_is_bf16)
    XB_CUDA_OK(launch(moe_fused_topk_kernel<__nv_bfloat16>, grid, block, 0, (cudaStream_t)stream, true, topk_weights, topk_ids,
                      reinterpret_cast<const __nv_bfloat16*>(gating_output), gating_stride, correction_bias, num_tokens,
                      num_experts, topk, renormalize, scoring_sigmoid));
  else
    XB_CUDA_OK(launch(moe_fused_topk_kernel<float>, grid, block, 0, (cudaStream_t)stream, true, topk_weights, topk_ids,
                      reinterpret_cast<const float*>(gating_output), gating_stride, correction_bias, num_tokens, num_experts,
                      topk, renormalize, scoring_sigmoid));
  return 0;
}

extern "C" int64_t xb_moe_experts_workspace_bytes(int num_tokens, int topk, int hidden, int inter) {
  return (int64_t)num_tokens * topk * ((int64_t)inter + hidden) * 2;
}

extern "C" int xb_moe_experts_bf16(void* out, int64_t out_stride, const void* input, int64_t in_stride,
                                   const int32_t* token_selected_experts, const float* token_final_scales,
                                   const void* fc1_weights, const void* fc2_weights, int num_tokens, int topk, int hidden,
                                   int inter, int num_local_experts, int expert_begin, void* workspace,
                                   int64_t workspace_bytes, xb_stream_t stream) {
  if (num_tokens == 0) return 0;
  XB_CHECK(hidden % 32 == 0 && inter % 32 == 0, "moe_experts: hidden %d / inter %d must be multiples of 32", hidden, inter);
  XB_CHECK(topk >= 1 && num_local_experts >= 1, "moe_experts: bad topk %d / experts %d", topk, num_local_experts);
  XB_CHECK((int64_t)num_tokens * topk <= 65535, "moe_experts: %d x %d (token, expert) pairs exceed the decode-sized path (65535)",
           num_tokens, topk);
  XB_CHECK(workspace != nullptr && workspace_bytes >= xb_moe_experts_workspace_bytes(num_tokens, topk, hidden, inter),
           "moe_experts: workspace too small (need %lld bytes)",
           (long long)xb_moe_experts_workspace_bytes(num_tokens, topk, hidden, inter));
  XB_CHECK(in_stride % 8 == 0 && out_stride % 2 == 0 &&
               ((reinterpret_cast<uintptr_t>(input) | reinterpret_cast<uintptr_t>(fc1_weights) |
                 reinterpret_cast<uintptr_t>(fc2_weights) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
           "moe_experts: 16-byte alignment");
  const int pairs = num_tokens * topk;
  auto* act = reinterpret_cast<__nv_bfloat16*>(workspace);                 // [pairs][inter]
  auto* y2 = act + (int64_t)pairs * inter;                                 // [pairs][hidden]
  cudaStream_t s = (cudaStream_t)stream;
  XB_CUDA_OK(launch(moe_expert_gemv_kernel<true>, dim3((inter + 15) / 16, pairs), dim3(256), 0, s, true, act, (int64_t)inter,
                    reinterpret_cast<const __nv_bfloat16*>(input), in_stride, topk,
                    reinterpret_cast<const __nv_bfloat16*>(fc1_weights), token_selected_experts, inter, hidden, expert_begin,
                    expert_begin + num_local_experts));
  XB_CUDA_OK(launch(moe_expert_gemv_kernel<false>, dim3((hidden + 15) / 16, pairs), dim3(256), 0, s, true, y2, (int64_t)hidden,
                    reinterpret_cast<const __nv_bfloat16*>(act), (int64_t)inter, 1,
                    reinterpret_cast<const __nv_bfloat16*>(fc2_weights), token_selected_experts, hidden, inter, expert_begin,
                    expert_begin + num_local_experts));
  const int cx = (hidden / 2 + 255) / 256;
  XB_CUDA_OK(launch(moe_combine_kernel, dim3(cx, num_tokens), dim3(256), 0, s, true, reinterpret_cast<__nv_bfloat16*>(out),
                    out_stride, reinterpret_cast<const __nv_bfloat16*>(y2), token_final_scales, topk, hidden));
  return 0;
}

extern "C" int xb_moe_experts_w4a16(void* out, int64_t out_stride, const void* input, int64_t in_stride,
                                    const int32_t* token_selected_experts, const float* token_final_scales,
                                    const uint32_t* fc1_qweight, const uint32_t* fc1_meta, const uint32_t* fc2_qweight,
                                    const uint32_t* fc2_meta, int group_size, int num_tokens, int topk, int hidden, int inter,
                                    int num_local_experts, int expert_begin, void* workspace, int64_t workspace_bytes,
                                    xb_stream_t stream) {
  if (num_tokens == 0) return 0;
  XB_CHECK(hidden % 64 == 0 && inter % 64 == 0, "moe_experts_w4a16: hidden %d / inter %d must be multiples of 64", hidden, inter);
  XB_CHECK(group_size >= 64 && group_size % 64 == 0 && hidden % group_size == 0 && inter % group_size == 0,
           "moe_experts_w4a16: group_size %d must be a multiple of 64 dividing hidden and inter", group_size);
  const int tpg = group_size / 64;
  XB_CHECK((tpg & (tpg - 1)) == 0, "moe_experts_w4a16: group_size / 64 must be a power of two");
  int gshift = 0;
  while ((1 << gshift) < tpg) ++gshift;
  XB_CHECK(topk >= 1 && num_local_experts >= 1 && (int64_t)num_tokens * topk <= 65535, "moe_experts_w4a16: bad topk / experts / pairs");
  XB_CHECK(workspace != nullptr && workspace_bytes >= xb_moe_experts_workspace_bytes(num_tokens, topk, hidden, inter),
           "moe_experts_w4a16: workspace too small (need %lld bytes)",
           (long long)xb_moe_experts_workspace_bytes(num_tokens, topk, hidden, inter));
  XB_CHECK(in_stride % 8 == 0 && out_stride % 2 == 0 &&
               ((reinterpret_cast<uintptr_t>(input) | reinterpret_cast<uintptr_t>(fc1_qweight) |
                 reinterpret_cast<uintptr_t>(fc2_qweight) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
           "moe_experts_w4a16: 16-byte alignment");
  const int pairs = num_tokens * topk;
  auto* act = reinterpret_cast<__nv_bfloat16*>(workspace);
  auto* y2 = act + (int64_t)pairs * inter;
  cudaStream_t s = (cudaStream_t)stream;
  XB_CUDA_OK(launch(moe_expert_gemv_w4_kernel<true>, dim3(inter / 16, pairs), dim3(256), 0, s, true, act, (int64_t)inter,
                    reinterpret_cast<const __nv_bfloat16*>(input), in_stride, topk, reinterpret_cast<const uint4*>(fc1_qweight),
                    fc1_meta, token_selected_experts, inter, hidden, gshift, expert_begin, expert_begin + num_local_experts));
  XB_CUDA_OK(launch(moe_expert_gemv_w4_kernel<false>, dim3(hidden / 16, pairs), dim3(256), 0, s, true, y2, (int64_t)hidden,
                    reinterpret_cast<const __nv_bfloat16*>(act), (int64_t)inter, 1, reinterpret_cast<const uint4*>(fc2_qweight),
                    fc2_meta, token_selected_experts, hidden, inter, gshift, expert_begin, expert_begin + num_local_experts));
  const int cx = (hidden / 2 + 255) / 256;
  XB_CUDA_OK(launch(moe_combine_kernel, dim3(cx, num_tokens), dim3(256), 0, s, true, reinterpret_cast<__nv_bfloat16*>(out),
                    out_stride, reinterpret_cast<const __nv_bfloat16*>(y2), token_final_scales, topk, hidden));
  return 0;
}
