// Step-boundary helpers of the decode path: embedding row gather
// (WordEmbeddingImpl::forward, xllm/core/layers/common/word_embedding_impl.cpp:33-56, TP=1 slice)
// and greedy argmax over the logits (the sampler's greedy branch).  Both are tiny and
// exist so that a whole decode step is made of this library's launches only.
#include "common.cuh"

namespace xb {

__global__ void __launch_bounds__(256)
embedding_kernel(__nv_bfloat16* __restrict__ out, const int32_t* __restrict__ token_ids,
                 const __nv_bfloat16* __restrict__ table, int hidden, int vocab) {
  pdl_launch_dependents();  // consumer prologues (weight / KV prefetch) may start now
  pdl_wait();
  const int64_t tok = blockIdx.x;
  int id = token_ids[tok];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const uint4* src = reinterpret_cast<const uint4*>(table + (int64_t)id * hidden);
  uint4* dst = reinterpret_cast<uint4*>(out + tok * hidden);
  for (int i = threadIdx.x; i < hidden / 8; i += blockDim.x) dst[i] = __ldg(src + i);
  pdl_launch_dependents();
}

// argmax over each row of logits[M, vocab] (bf16); ties -> lowest index (torch.argmax semantics on CUDA
// are unspecified for ties; lowest index is what the CPU oracle's torch.argmax returns).
__global__ void __launch_bounds__(1024)
argmax_kernel(int32_t* __restrict__ out, const __nv_bfloat16* __restrict__ logits, int64_t stride, int vocab) {
  __shared__ float sv[32];
  __shared__ int si[32];
  pdl_launch_dependents();  // consumer prologues (weight / KV prefetch) may start now
  pdl_wait();
  const __nv_bfloat16* row = logits + (int64_t)blockIdx.x * stride;
  float best = -INFINITY;
  int besti = 0x7fffffff;
  const int nvec = vocab / 8;
  const uint4* rv = reinterpret_cast<const uint4*>(row);
  // one CTA per row: batches of 8 independent 16-byte loads per thread (the row is read in 3 round trips for a 152 K
  // vocabulary instead of 19 dependent ones)
  constexpr int kBatch = 8;
  for (int i0 = threadIdx.x; i0 < nvec; i0 += blockDim.x * kBatch) {
    uint4 v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = i0 + u * blockDim.x;
      v[u] = i < nvec ? __ldcs(rv + i) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i >= nvec) continue;
      const uint32_t* p = &v[u].x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = bf16lo(p[j]), b = bf16hi(p[j]);
        int ia = i * 8 + 2 * j, ib = ia + 1;
        if (a > best || (a == best && ia < besti)) { best = a; besti = ia; }
        if (b > best || (b == best && ib < besti)) { best = b; besti = ib; }
      }
    }
  }
  for (int i = nvec * 8 + threadIdx.x; i < vocab; i += blockDim.x) {
    float a = __bfloat162float(row[i]);
    if (a > best || (a == best && i < besti)) { best = a; besti = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = besti; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
    besti = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if (threadIdx.x == 0) out[blockIdx.x] = besti;
  }
  pdl_launch_dependents();
}

// Decode-step metadata refresh of a CUDA-graph replay, on the device (one launch, no host sync): copies the live step's
// tokens / positions / slots / paged triplet into the graph's persistent buffers, zero-pads the token rows up to the
// captured batch and derives the per-request kv lengths.  Same field-by-field semantics as
// xllm::kernel::cuda::llm_decode_metadata_update_kernel (kernels/cuda/llm_decode_metadata_update.cu:29-62), which
// CudaGraphPersistentParam::update_llm_decode_metadata_fast_path launches before every replay
// (runtime/cuda_graph_executor_impl.cpp:218-258).  Plus, optionally, what the reference still does on the HOST before each
// replay - re-running the attention `plan` for the new context lengths (cuda_graph_executor_impl.cpp:751-822): the split
// geometry of this library's kernels is derived on the device from the refreshed kv_indptr / last_page_len, so the only
// per-replay state a plan owns are the arrival counters, which this kernel re-zeroes (n_counter_words > 0).
struct DecodeMetaParams {
  const int32_t* src_tokens;
  const int32_t* src_positions;
  const int32_t* src_new_cache_slots;
  const int32_t* src_kv_seq_lens;
  const int32_t* src_paged_kv_indptr;
  const int32_t* src_paged_kv_indices;
  const int32_t* src_paged_kv_last_page_len;
  int32_t* dst_tokens;
  int32_t* dst_positions;
  int32_t* dst_new_cache_slots;
  int32_t* dst_kv_seq_lens;
  int32_t* dst_kv_seq_lens_delta;
  int32_t* dst_paged_kv_indptr;
  int32_t* dst_paged_kv_indices;
  int32_t* dst_paged_kv_last_page_len;
  int64_t actual_num_tokens, padded_num_tokens, actual_batch_size, actual_indices_size;
  int32_t* plan_counters;        // int workspace of the decode plan (arrival counters), or null
  int64_t n_counter_words;
};

__global__ void __launch_bounds__(256)
decode_metadata_update_kernel(const DecodeMetaParams p, int64_t max_work) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t step = (int64_t)blockDim.x * gridDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < max_work; i += step) {
    if (i < p.actual_num_tokens) {
      p.dst_tokens[i] = p.src_tokens[i];
      p.dst_positions[i] = p.src_positions[i];
      p.dst_new_cache_slots[i] = p.src_new_cache_slots[i];
    } else if (i < p.padded_num_tokens) {       // padding rows of the captured batch: token 0, slot 0 (block 0 = padding block)
      p.dst_tokens[i] = 0;
      p.dst_new_cache_slots[i] = 0;
    }
    if (i < p.actual_batch_size + 1) {
      p.dst_kv_seq_lens[i] = p.src_kv_seq_lens[i];
      p.dst_paged_kv_indptr[i] = p.src_paged_kv_indptr[i];
    }
    if (i < p.actual_batch_size) {
      p.dst_kv_seq_lens_delta[i] = p.src_kv_seq_lens[i + 1] - p.src_kv_seq_lens[i];
      p.dst_paged_kv_last_page_len[i] = p.src_paged_kv_last_page_len[i];
    }
    if (i < p.actual_indices_size) p.dst_paged_kv_indices[i] = p.src_paged_kv_indices[i];
    if (i < p.n_counter_words) p.plan_counters[i] = 0;
  }
}

}  // namespace xb

using namespace xb;

extern "C" int xb_decode_metadata_update(const int32_t* src_tokens, const int32_t* src_positions,
                                         const int32_t* src_new_cache_slots, const int32_t* src_kv_seq_lens,
                                         const int32_t* src_paged_kv_indptr, const int32_t* src_paged_kv_indices,
                                         const int32_t* src_paged_kv_last_page_len, int32_t* dst_tokens,
                                         int32_t* dst_positions, int32_t* dst_new_cache_slots, int32_t* dst_kv_seq_lens,
                                         int32_t* dst_kv_seq_lens_delta, int32_t* dst_paged_kv_indptr,
                                         int32_t* dst_paged_kv_indices, int32_t* dst_paged_kv_last_page_len,
                                         int64_t actual_num_tokens, int64_t padded_num_tokens, int64_t actual_batch_size,
                                         int64_t actual_indices_size, int32_t* plan_counters, int64_t n_counter_words,
                                         xb_stream_t stream) {
  XB_CHECK(actual_num_tokens >= 0 && padded_num_tokens >= 0 && actual_batch_size >= 0 && actual_indices_size >= 0 &&
               n_counter_words >= 0,
           "decode_metadata_update: negative size");
  XB_CHECK(n_counter_words == 0 || plan_counters != nullptr, "decode_metadata_update: plan_counters is null");
  DecodeMetaParams p{src_tokens, src_positions, src_new_cache_slots, src_kv_seq_lens, src_paged_kv_indptr,
                     src_paged_kv_indices, src_paged_kv_last_page_len, dst_tokens, dst_positions, dst_new_cache_slots,
                     dst_kv_seq_lens, dst_kv_seq_lens_delta, dst_paged_kv_indptr, dst_paged_kv_indices,
                     dst_paged_kv_last_page_len, actual_num_tokens, padded_num_tokens, actual_batch_size,
                     actual_indices_size, plan_counters, n_counter_words};
  int64_t max_work = actual_num_tokens;
  if (padded_num_tokens > max_work) max_work = padded_num_tokens;
  if (actual_batch_size + 1 > max_work) max_work = actual_batch_size + 1;
  if (actual_indices_size > max_work) max_work = actual_indices_size;
  if (n_counter_words > max_work) max_work = n_counter_words;
  if (max_work <= 0) return 0;
  int64_t blocks = (max_work + 255) / 256;
  if (blocks > 4096) blocks = 4096;            // strided loop: bounded launch size (llm_decode_metadata_update.cu:84-87)
  XB_CUDA_OK(launch(decode_metadata_update_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, true, p, max_work));
  return 0;
}

extern "C" int xb_embedding_bf16(void* out, const int32_t* token_ids, const void* table, int num_tokens, int hidden,
                                 int vocab, xb_stream_t stream) {
  if (num_tokens == 0) return 0;
  XB_CHECK(hidden % 8 == 0, "embedding: hidden %d must be a multiple of 8", hidden);
  XB_CUDA_OK(launch(embedding_kernel, dim3(num_tokens), dim3(256), 0, (cudaStream_t)stream, true,
                    reinterpret_cast<__nv_bfloat16*>(out), token_ids, reinterpret_cast<const __nv_bfloat16*>(table),
                    hidden, vocab));
  return 0;
}

extern "C" int xb_argmax_bf16(int32_t* out, const void* logits, int64_t stride, int rows, int vocab,
                              xb_stream_t stream) {
  if (rows == 0) return 0;
  XB_CHECK(stride % 8 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0, "argmax: logits must be 16B aligned");
  XB_CUDA_OK(launch(argmax_kernel, dim3(rows), dim3(1024), 0, (cudaStream_t)stream, true, out,
                    reinterpret_cast<const __nv_bfloat16*>(logits), stride, vocab));
  return 0;
}
