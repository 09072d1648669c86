// Test-side Python binding of the link-time drop-in (csrc/shim/xllm_cuda_ops.cpp): the xllm::kernel::cuda::* functions, called
// with torch tensors exactly as xLLM's layers call them, so that tests can exercise the C++ boundary itself - argument checks
// that raise c10::Error (the reference's tests rely on that: tests/core/kernels/cuda/cutlass_scaled_mm_test.cpp:279-295) and,
// on a GPU, the same launches as the ctypes driver.  Not part of the product; built by xllm_b200/build_shim.py::build_py().
#include <torch/extension.h>

#include <optional>
#include <string>
#include <tuple>

namespace xllm::kernel::cuda {
// declarations as in cuda_ops_api.h:31-266 (defined in xllm_cuda_ops.cpp)
void rotary_embedding(torch::Tensor& positions, torch::Tensor& query, std::optional<torch::Tensor> key, torch::Tensor& cos_sin_cache,
                      bool is_neox);
void act_and_mul(torch::Tensor out, torch::Tensor input, const std::string& act_mode);
void reshape_paged_cache(torch::Tensor slot_ids, torch::Tensor keys, torch::Tensor values, torch::Tensor key_cache,
                         torch::Tensor value_cache);
void rms_norm(torch::Tensor output, torch::Tensor input, torch::Tensor weight, double eps);
void fused_add_rms_norm(torch::Tensor& input, torch::Tensor& residual, torch::Tensor& weight, double epsilon);
torch::Tensor matmul(torch::Tensor a, torch::Tensor b, std::optional<torch::Tensor> bias);
void cutlass_scaled_mm(torch::Tensor& c, torch::Tensor const& a, torch::Tensor const& b, torch::Tensor const& a_scales,
                       torch::Tensor const& b_scales, std::optional<torch::Tensor> const& bias);
void static_scaled_fp8_quant(torch::Tensor& out, torch::Tensor const& input, torch::Tensor const& scale);
std::tuple<torch::Tensor, torch::Tensor> fp8_scaled_quantize(const torch::Tensor& input, const std::optional<torch::Tensor>& output,
                                                             const std::optional<torch::Tensor>& scale);
void rms_norm_static_fp8_quant(torch::Tensor& out, torch::Tensor& input, torch::Tensor& weight, torch::Tensor& scale, double epsilon);
void fused_add_rms_norm_static_fp8_quant(torch::Tensor& out, torch::Tensor& input, torch::Tensor& residual, torch::Tensor& weight,
                                         torch::Tensor& scale, double epsilon);
void fused_qk_norm_rope(torch::Tensor& qkv, int64_t num_heads_q, int64_t num_heads_k, int64_t num_heads_v, int64_t head_dim, double eps,
                        const torch::Tensor& q_weight, const torch::Tensor& k_weight, const torch::Tensor& cos_sin_cache,
                        bool interleaved, const torch::Tensor& position_ids);
std::tuple<torch::Tensor, torch::Tensor> moe_fused_topk(torch::Tensor& gating_output, int64_t topk, bool renormalize,
                                                        const std::optional<torch::Tensor>& correction_bias,
                                                        const std::string& scoring_func);
}  // namespace xllm::kernel::cuda

namespace xk = xllm::kernel::cuda;

PYBIND11_MODULE(xllm_b200_shim_py, m) {
  m.doc() = "xllm::kernel::cuda::* of libxllm_b200_shim.so (test binding)";
  m.def("rotary_embedding", [](torch::Tensor positions, torch::Tensor query, std::optional<torch::Tensor> key,
                               torch::Tensor cos_sin_cache, bool is_neox) { xk::rotary_embedding(positions, query, key, cos_sin_cache, is_neox); });
  m.def("act_and_mul", [](torch::Tensor out, torch::Tensor input, const std::string& mode) { xk::act_and_mul(out, input, mode); });
  m.def("reshape_paged_cache", [](torch::Tensor slot_ids, torch::Tensor keys, torch::Tensor values, torch::Tensor key_cache,
                                  torch::Tensor value_cache) { xk::reshape_paged_cache(slot_ids, keys, values, key_cache, value_cache); });
  m.def("rms_norm", [](torch::Tensor output, torch::Tensor input, torch::Tensor weight, double eps) { xk::rms_norm(output, input, weight, eps); });
  m.def("fused_add_rms_norm", [](torch::Tensor input, torch::Tensor residual, torch::Tensor weight, double eps) {
    xk::fused_add_rms_norm(input, residual, weight, eps);
  });
  m.def("matmul", [](torch::Tensor a, torch::Tensor b, std::optional<torch::Tensor> bias) { return xk::matmul(a, b, bias); });
  m.def("cutlass_scaled_mm", [](torch::Tensor c, torch::Tensor a, torch::Tensor b, torch::Tensor a_scales, torch::Tensor b_scales,
                                std::optional<torch::Tensor> bias) { xk::cutlass_scaled_mm(c, a, b, a_scales, b_scales, bias); });
  m.def("static_scaled_fp8_quant", [](torch::Tensor out, torch::Tensor input, torch::Tensor scale) { xk::static_scaled_fp8_quant(out, input, scale); });
  m.def("fp8_scaled_quantize", [](torch::Tensor input, std::optional<torch::Tensor> output, std::optional<torch::Tensor> scale) {
    return xk::fp8_scaled_quantize(input, output, scale);
  });
  m.def("rms_norm_static_fp8_quant", [](torch::Tensor out, torch::Tensor input, torch::Tensor weight, torch::Tensor scale, double eps) {
    xk::rms_norm_static_fp8_quant(out, input, weight, scale, eps);
  });
  m.def("fused_add_rms_norm_static_fp8_quant", [](torch::Tensor out, torch::Tensor input, torch::Tensor residual, torch::Tensor weight,
                                                  torch::Tensor scale, double eps) {
    xk::fused_add_rms_norm_static_fp8_quant(out, input, residual, weight, scale, eps);
  });
  m.def("fused_qk_norm_rope", [](torch::Tensor qkv, int64_t hq, int64_t hk, int64_t hv, int64_t head_dim, double eps, torch::Tensor qw,
                                 torch::Tensor kw, torch::Tensor cos_sin, bool interleaved, torch::Tensor pos) {
    xk::fused_qk_norm_rope(qkv, hq, hk, hv, head_dim, eps, qw, kw, cos_sin, interleaved, pos);
  });
  m.def("moe_fused_topk", [](torch::Tensor gating, int64_t topk, bool renormalize, std::optional<torch::Tensor> bias,
                             const std::string& scoring) { return xk::moe_fused_topk(gating, topk, renormalize, bias, scoring); });
}
