// Python binding of the REFERENCE's own CUDA kernels (xllm::kernel::cuda::* compiled from /root/reference/xllm/core/kernels/cuda
// by oracle/build_ref.py into oracle/_ref/): the functions are declared here exactly as cuda_ops_api.h:31-221 declares them and
// defined by the reference's translation units.  TEST INFRASTRUCTURE: lets the GPU parity tests compare this library's kernels with
// the reference's kernels on the same inputs.  Never loaded by the product.
#include <torch/extension.h>

#include <optional>
#include <string>

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
  m.def("fused_qk_norm_rope", [](torch::Tensor qkv, int64_t hq, int64_t hk, int64_t hv, int64_t d, double eps, torch::Tensor qw, torch::Tensor kw,
                                 torch::Tensor cache, bool interleaved, torch::Tensor pos) {
    xk::fused_qk_norm_rope(qkv, hq, hk, hv, d, eps, qw, kw, cache, interleaved, pos);
  });
}
