"""Mixture-of-experts restatement.  TEST INFRASTRUCTURE - see oracle/__init__.py.

Router: xllm::kernel::cuda::moe_fused_topk (xllm/core/kernels/cuda/moe/moe_fused_topk.cu:22-58):
  softmax  - moe_topk_softmax_kernels.cuh: fp32 softmax over the experts, k rounds of arg-max (ties -> lower expert index,
             the ordering tests/core/kernels/cuda/moe/moe_topk_test.cu:31-55 pins), weights = the selected probabilities;
             the correction bias is NOT used on this path (moe_fused_topk.cu:36-43 passes nullopt)
  sigmoid  - moe_topk_sigmoid_kernels.cuh:48-72,79-150: v = 1 / (1 + expf(-x)) (+ bias[e]) is the SELECTION score, the
             returned weight is v_selected - bias[e] (a float subtraction of the biased value, not the un-biased sigmoid)
  renormalize: weights * (1 / sum of the k selected weights), the sum taken in selection order (:144-150).
Experts: xllm::kernel::cuda::cutlass_fused_moe (moe/fused_moe.cpp:23-124) runs FlashInfer's fused_moe_100 (TRT-LLM grouped GEMM,
not vendored: "parity unpinned"); restated from its published structure for unquantised bf16 experts:
  fc1 [E, 2I, H] rows in [up | gate] order (layers/cuda/fused_moe.cpp:124-126), SwiGLU on the fp32 GEMM1 sums, one rounding
  to bf16; GEMM2 to bf16; finalize: out = bf16(sum_k scale_k * y2_k) in fp32, selection order.
"""
import numpy as np
import torch

BF16 = torch.bfloat16
F32 = torch.float32


def moe_fused_topk(gating_output: torch.Tensor, topk: int, renormalize: bool, correction_bias=None, scoring_func="softmax"):
    g = gating_output.to(F32).numpy().astype(np.float32)
    T, E = g.shape
    if scoring_func == "softmax":
        m = g.max(axis=1, keepdims=True)
        e = np.exp((g - m).astype(np.float32)).astype(np.float32)
        sel = (e * (np.float32(1.0) / e.sum(axis=1, keepdims=True, dtype=np.float32))).astype(np.float32)
        bias = None
    elif scoring_func == "sigmoid":
        sel = (np.float32(1.0) / (np.float32(1.0) + np.exp(-g).astype(np.float32))).astype(np.float32)
        bias = correction_bias.to(F32).numpy().astype(np.float32) if correction_bias is not None else None
        if bias is not None:
            sel = (sel + bias[None, :]).astype(np.float32)
    else:
        raise ValueError(f"Unsupported scoring function for moe topk: {scoring_func}")
    w = np.zeros((T, topk), np.float32)
    ids = np.zeros((T, topk), np.int32)
    for t in range(T):
        cand = sel[t].copy()
        row_sum = np.float32(0)
        for k in range(topk):
            best = int(np.flatnonzero(cand == cand.max())[0])       # ties -> lowest expert index
            v = np.float32(cand[best])
            if bias is not None:
                v = np.float32(v - bias[best])
            w[t, k], ids[t, k] = v, best
            row_sum = np.float32(row_sum + v)
            cand[best] = -np.inf
        if renormalize:
            w[t] = (w[t] * np.float32(np.float32(1.0) / row_sum)).astype(np.float32)
    return torch.from_numpy(w), torch.from_numpy(ids)


def fused_moe(x, token_selected_experts, token_final_scales, fc1, fc2, expert_begin=0, return_abs=False):
    """x [T, H] bf16, ids int32 [T, k], scales fp32 [T, k], fc1 [E_local, 2I, H] ([up | gate]), fc2 [E_local, H, I] -> [T, H]
    bf16.  Experts outside [expert_begin, expert_begin + E_local) contribute zero."""
    T, H = x.shape
    El, I2, _ = fc1.shape
    inter = I2 // 2
    out = torch.zeros(T, H, dtype=F32)
    mag = torch.zeros(T, H, dtype=F32)          # sum_k scale_k |y2_k|: the magnitude the rounding errors of the terms scale with
    for t in range(T):
        acc = torch.zeros(H, dtype=F32)
        for k in range(token_selected_experts.size(1)):
            e = int(token_selected_experts[t, k]) - expert_begin
            if 0 <= e < El:
                h1 = fc1[e].to(F32) @ x[t].to(F32)
                up, gate = h1[:inter], h1[inter:]
                a = ((gate / (1.0 + torch.exp(-gate))) * up).to(BF16)
                y2 = (fc2[e].to(F32) @ a.to(F32)).to(BF16).to(F32)
            else:
                y2 = torch.zeros(H, dtype=F32)
            acc = torch.addcmul(acc, y2, token_final_scales[t, k].to(F32))
            mag[t] += y2.abs() * token_final_scales[t, k].to(F32).abs()
        out[t] = acc
    return (out.to(BF16), mag) if return_abs else out.to(BF16)
