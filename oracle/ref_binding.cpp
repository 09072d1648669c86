// Python binding of the REFERENCE's own CUDA kernels (xllm::kernel::cuda::* compiled from /root/reference/xllm/core/kernels/cuda
// by oracle/build_ref.py into oracle/_ref/): the functions are declared here exactly as cuda_ops_api.h:31-221 declares them and
// defined by the reference's translation units.  TEST INFRASTRUCTURE: lets the GPU parity tests compare this library's kernels with
// the reference's kernels on the same inputs.  Never loaded by the product.
#include <ATen/cuda/CUDAContext.h>
#include <cuda_runtime.h>
#include <torch/extension.h>

#include <optional>
#include <string>
#include <tuple>
#include <vector>

namespace xllm::kernel::cuda {
void rotary_embedding(torch::Tensor& positions, torch::Tensor& query, std::optional<torch::Tensor> key, torch::Tensor& cos_sin_cache,
                      bool is_neox);
void act_and_mul(torch::Tensor out, torch::Tensor input, const std::string& act_mode);
void reshape_paged_cache(torch::Tensor slot_ids, torch::Tensor keys, torch::Tensor values, torch::Tensor key_cache,
                         torch::Tensor value_cache);
void rms_norm(torch::Tensor output, torch::Tensor input, torch::Tensor weight, double eps);
void fused_add_rms_norm(torch::Tensor& input, torch::Tensor& residual, torch::Tensor& weight, double epsilon);
void static_scaled_fp8_quant(torch::Tensor& out, torch::Tensor const& input, torch::Tensor const& scale);
void rms_norm_static_fp8_quant(torch::Tensor& out, torch::Tensor& input, torch::Tensor& weight, torch::Tensor& scale, double epsilon);
void fused_add_rms_norm_static_fp8_quant(torch::Tensor& out, torch::Tensor& input, torch::Tensor& residual, torch::Tensor& weight,
                                         torch::Tensor& scale, double epsilon);
void fused_qk_norm_rope(torch::Tensor& qkv, int64_t num_heads_q, int64_t num_heads_k, int64_t num_heads_v, int64_t head_dim, double eps,
                        const torch::Tensor& q_weight, const torch::Tensor& k_weight, const torch::Tensor& cos_sin_cache,
                        bool interleaved, const torch::Tensor& position_ids);
std::tuple<torch::Tensor, torch::Tensor> fp8_scaled_quantize(const torch::Tensor& input, const std::optional<torch::Tensor>& output,
                                                             const std::optional<torch::Tensor>& scale);
std::tuple<torch::Tensor, torch::Tensor> moe_fused_topk(torch::Tensor& gating_output, int64_t topk, bool renormalize,
                                                        const std::optional<torch::Tensor>& correction_bias,
                                                        const std::string& scoring_func);
// llm_decode_metadata_update.h:35-58
struct LlmDecodeMetadataUpdateParams {
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
  int64_t actual_num_tokens;
  int64_t padded_num_tokens;
  int64_t actual_batch_size;
  int64_t actual_indices_size;
};
void update_llm_decode_metadata(const LlmDecodeMetadataUpdateParams& params, cudaStream_t stream);
}  // namespace xllm::kernel::cuda

namespace xk = xllm::kernel::cuda;

PYBIND11_MODULE(xllm_ref_kernels_py, m) {
  m.doc() = "the reference's xllm::kernel::cuda::* kernels, compiled from its own sources (oracle/_ref)";
  m.def("rotary_embedding", [](torch::Tensor positions, torch::Tensor query, std::optional<torch::Tensor> key, torch::Tensor cache,
                               bool is_neox) { xk::rotary_embedding(positions, query, key, cache, is_neox); });
  m.def("act_and_mul", [](torch::Tensor out, torch::Tensor input, const std::string& mode) { xk::act_and_mul(out, input, mode); });
  m.def("reshape_paged_cache", [](torch::Tensor slots, torch::Tensor k, torch::Tensor v, torch::Tensor kc, torch::Tensor vc) {
    xk::reshape_paged_cache(slots, k, v, kc, vc);
  });
  m.def("rms_norm", [](torch::Tensor out, torch::Tensor in, torch::Tensor w, double eps) { xk::rms_norm(out, in, w, eps); });
  m.def("fused_add_rms_norm", [](torch::Tensor in, torch::Tensor res, torch::Tensor w, double eps) { xk::fused_add_rms_norm(in, res, w, eps); });
  m.def("static_scaled_fp8_quant", [](torch::Tensor out, torch::Tensor in, torch::Tensor scale) { xk::static_scaled_fp8_quant(out, in, scale); });
  m.def("rms_norm_static_fp8_quant", [](torch::Tensor out, torch::Tensor in, torch::Tensor w, torch::Tensor scale, double eps) {
    xk::rms_norm_static_fp8_quant(out, in, w, scale, eps);
  });
  m.def("fused_add_rms_norm_static_fp8_quant", [](torch::Tensor out, torch::Tensor in, torch::Tensor res, torch::Tensor w, torch::Tensor scale,
                                                  double eps) { xk::fused_add_rms_norm_static_fp8_quant(out, in, res, w, scale, eps); });
  m.def("fp8_scaled_quantize", [](torch::Tensor input, std::optional<torch::Tensor> output, std::optional<torch::Tensor> scale) {
    return xk::fp8_scaled_quantize(input, output, scale);
  });
  m.def("moe_fused_topk", [](torch::Tensor gating, int64_t topk, bool renormalize, std::optional<torch::Tensor> bias,
                             const std::string& scoring) { return xk::moe_fused_topk(gating, topk, renormalize, bias, scoring); });
  // src: tokens, positions, new_cache_slots, kv_seq_lens, paged_kv_indptr, paged_kv_indices, paged_kv_last_page_len (int32 CUDA);
  // dst: the same seven + kv_seq_lens_delta inserted after kv_seq_lens (the struct's field order)
  m.def("update_llm_decode_metadata", [](std::vector<torch::Tensor> src, std::vector<torch::Tensor> dst, int64_t actual_num_tokens,
                                         int64_t padded_num_tokens, int64_t actual_batch_size, int64_t actual_indices_size) {
    TORCH_CHECK(src.size() == 7 && dst.size() == 8, "7 source and 8 destination tensors expected");
    for (auto& t : src) TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kInt32 && t.is_contiguous());
    for (auto& t : dst) TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kInt32 && t.is_contiguous());
    xk::LlmDecodeMetadataUpdateParams p{};
    p.src_tokens = src[0].data_ptr<int32_t>();
    p.src_positions = src[1].data_ptr<int32_t>();
    p.src_new_cache_slots = src[2].data_ptr<int32_t>();
    p.src_kv_seq_lens = src[3].data_ptr<int32_t>();
    p.src_paged_kv_indptr = src[4].data_ptr<int32_t>();
    p.src_paged_kv_indices = src[5].data_ptr<int32_t>();
    p.src_paged_kv_last_page_len = src[6].data_ptr<int32_t>();
    p.dst_tokens = dst[0].data_ptr<int32_t>();
    p.dst_positions = dst[1].data_ptr<int32_t>();
    p.dst_new_cache_slots = dst[2].data_ptr<int32_t>();
    p.dst_kv_seq_lens = dst[3].data_ptr<int32_t>();
    p.dst_kv_seq_lens_delta = dst[4].data_ptr<int32_t>();
    p.dst_paged_kv_indptr = dst[5].data_ptr<int32_t>();
    p.dst_paged_kv_indices = dst[6].data_ptr<int32_t>();
    p.dst_paged_kv_last_page_len = dst[7].data_ptr<int32_t>();
    p.actual_num_tokens = actual_num_tokens;
    p.padded_num_tokens = padded_num_tokens;
    p.actual_batch_size = actual_batch_size;
    p.actual_indices_size = actual_indices_size;
    xk::update_llm_decode_metadata(p, at::cuda::getCurrentCUDAStream().stream());
  });
  m.def("fused_qk_norm_rope", [](torch::Tensor qkv, int64_t hq, int64_t hk, int64_t hv, int64_t d, double eps, torch::Tensor qw, torch::Tensor kw,
                                 torch::Tensor cache, bool interleaved, torch::Tensor pos) {
    xk::fused_qk_norm_rope(qkv, hq, hk, hv, d, eps, qw, kw, cache, interleaved, pos);
  });
}
