"""The C++ link-time boundary (csrc/shim/xllm_cuda_ops.cpp) called as xLLM's layers call it - torch tensors through the
xllm::kernel::cuda::* signatures - via the test binding csrc/shim/shim_py.cpp.  On CPU only the argument checks can run: they
must raise c10::Error (RuntimeError in Python) before any CUDA call, as the reference's tests expect
(tests/core/kernels/cuda/cutlass_scaled_mm_test.cpp:274-295), and never reinterpret a tensor of the wrong dtype or device."""
import pytest
import torch


@pytest.fixture(scope="module")
def shim(built_lib):
    from xllm_b200 import build_shim
    return build_shim.load_py()


def test_cutlass_scaled_mm_rejects_mismatched_dimensions(shim):
    # the reference's TestInvalidInputs case: b is [K + 1, N]
    M, N, K = 64, 128, 256
    a = torch.randn(M, K).to(torch.float8_e4m3fn)
    b_wrong = torch.randn(K + 1, N).to(torch.float8_e4m3fn).t().contiguous().t()
    c = torch.zeros(M, N, dtype=torch.bfloat16)
    one = torch.ones(1)
    with pytest.raises(RuntimeError):
        shim.cutlass_scaled_mm(c, a, b_wrong, one, one, None)
    with pytest.raises(RuntimeError):                                   # row-major b: the reference requires column-major
        shim.cutlass_scaled_mm(c, a, torch.randn(K, N).to(torch.float8_e4m3fn), one, one, None)
    with pytest.raises(RuntimeError):                                   # 3-D input
        shim.cutlass_scaled_mm(c, a.view(1, M, K), b_wrong, one, one, None)


def test_unsupported_modes_raise_with_the_reference_messages(shim):
    x = torch.zeros(2, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="Unsupported act mode"):       # activation.cu:172-186
        shim.act_and_mul(torch.zeros(2, 4, dtype=torch.bfloat16), x, "relu")
    with pytest.raises(RuntimeError, match="Unsupported scoring function"):   # moe_fused_topk.cu:52-56
        shim.moe_fused_topk(torch.zeros(2, 8), 2, True, None, "tanh")


def test_cpu_tensors_and_other_dtypes_are_rejected_not_reinterpreted(shim):
    bf = lambda *s: torch.zeros(*s, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA bfloat16"):
        shim.rms_norm(bf(2, 8), bf(2, 8), bf(8), 1e-6)
    with pytest.raises(RuntimeError, match="CUDA bfloat16"):
        shim.fused_add_rms_norm(bf(2, 8), bf(2, 8), bf(8), 1e-6)
    with pytest.raises(RuntimeError, match="CUDA bfloat16"):
        shim.matmul(torch.zeros(2, 8, dtype=torch.float16), bf(4, 8), None)
    with pytest.raises(RuntimeError, match="CUDA bfloat16"):
        shim.rotary_embedding(torch.zeros(2, dtype=torch.int64), bf(2, 16), None, bf(32, 8), True)
    with pytest.raises(RuntimeError, match="CUDA bfloat16"):
        shim.act_and_mul(bf(2, 4), bf(2, 8), "silu")
    with pytest.raises(RuntimeError):
        shim.static_scaled_fp8_quant(torch.zeros(2, 8, dtype=torch.float8_e4m3fn), bf(2, 8), torch.ones(1))
    with pytest.raises(RuntimeError):
        shim.fp8_scaled_quantize(bf(2, 8), None, None)
    with pytest.raises(RuntimeError):
        shim.reshape_paged_cache(torch.zeros(2, dtype=torch.int32), bf(2, 1, 8), bf(2, 1, 8), bf(4, 2, 1, 8), bf(4, 2, 1, 8))
