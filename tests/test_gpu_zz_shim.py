"""The C++ link-time boundary on the GPU: xllm::kernel::cuda::* of csrc/shim/xllm_cuda_ops.cpp called with torch tensors (test
binding csrc/shim/shim_py.cpp), as xLLM's layers call the reference's functions.  It must launch exactly what the ctypes driver
(xllm_b200/ops.py) launches - same C ABI underneath - so results are compared bit for bit with it (the driver itself is checked
against the oracle everywhere else), plus the reference's own invalid-input cases on CUDA tensors
(tests/core/kernels/cuda/cutlass_scaled_mm_test.cpp:274-295)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV, BF16, E4M3 = "cuda", torch.bfloat16, torch.float8_e4m3fn


@pytest.fixture(scope="module")
def shim(built_lib):
    from xllm_b200 import build_shim, ops
    yield build_shim.load_py()
    # the shim's cutlass_scaled_mm registers its own split-K workspace with the library on first use: withdraw it so that tests
    # collected after this module see the library's default state
    torch.cuda.synchronize()
    ops.disable_fp8_splitk()


def _g(seed):
    return torch.Generator(device=DEV).manual_seed(seed)


def test_norms_and_activation_through_the_cpp_boundary(shim):
    from xllm_b200 import ops
    g = _g(1)
    x = torch.randn(7, 3584, generator=g, device=DEV).to(BF16)
    r = torch.randn(7, 3584, generator=g, device=DEV).to(BF16)
    w = (1 + 0.1 * torch.randn(3584, generator=g, device=DEV)).to(BF16)
    a, b = torch.empty_like(x), torch.empty_like(x)
    shim.rms_norm(a, x, w, 1e-6)
    ops.rms_norm(b, x, w, 1e-6)
    assert torch.equal(a, b)
    x1, r1, x2, r2 = x.clone(), r.clone(), x.clone(), r.clone()
    shim.fused_add_rms_norm(x1, r1, w, 1e-6)
    ops.fused_add_rms_norm(x2, r2, w, 1e-6)
    assert torch.equal(x1, x2) and torch.equal(r1, r2)
    gu = torch.randn(7, 2 * 1024, generator=g, device=DEV).to(BF16)
    o1, o2 = torch.empty(7, 1024, dtype=BF16, device=DEV), torch.empty(7, 1024, dtype=BF16, device=DEV)
    shim.act_and_mul(o1, gu, "silu")
    ops.act_and_mul(o2, gu, "silu")
    assert torch.equal(o1, o2)
    s = torch.tensor([0.05], device=DEV)
    q1, q2 = torch.empty(7, 3584, dtype=E4M3, device=DEV), torch.empty(7, 3584, dtype=E4M3, device=DEV)
    shim.rms_norm_static_fp8_quant(q1, x, w, s, 1e-6)
    ops.rms_norm_static_fp8_quant(q2, x, w, s, 1e-6)
    assert torch.equal(q1.view(torch.uint8), q2.view(torch.uint8))


@pytest.mark.parametrize("M", [4, 300])
def test_matmul_through_the_cpp_boundary(M, shim):
    from xllm_b200 import ops
    g = _g(2)
    a = torch.randn(M, 1024, generator=g, device=DEV).to(BF16)
    w = (torch.randn(512, 1024, generator=g, device=DEV) * 0.05).to(BF16)
    bias = torch.randn(512, generator=g, device=DEV).to(BF16)
    y = shim.matmul(a, w, bias)
    ref = torch.empty(M, 512, dtype=BF16, device=DEV)
    ops.matmul(a, w, bias, ref)
    assert y.shape == (M, 512) and torch.equal(y, ref)


def test_cutlass_scaled_mm_through_the_cpp_boundary_and_its_invalid_inputs(shim):
    from xllm_b200 import ops
    g = _g(3)
    M, N, K = 64, 128, 256                                              # the reference test's sizes
    a = torch.randn(M, K, generator=g, device=DEV).to(E4M3)
    b = torch.randn(K, N, generator=g, device=DEV).to(E4M3).t().contiguous().t()     # column-major [K, N]
    one = torch.ones(1, device=DEV)
    c, ref = torch.zeros(M, N, dtype=BF16, device=DEV), torch.zeros(M, N, dtype=BF16, device=DEV)
    shim.cutlass_scaled_mm(c, a, b, one, one, None)
    ops.cutlass_scaled_mm(ref, a, b, one, one, None)
    torch.cuda.synchronize()
    assert torch.equal(c, ref)
    b_wrong = torch.randn(K + 1, N, generator=g, device=DEV).to(E4M3).t().contiguous().t()
    with pytest.raises(RuntimeError):
        shim.cutlass_scaled_mm(c, a, b_wrong, one, one, None)
    with pytest.raises(RuntimeError):
        shim.cutlass_scaled_mm(c, a, b, one, one, torch.randn(N + 1, device=DEV).to(BF16))
