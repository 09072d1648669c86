#!/bin/bash
# One gpurun call: first run of the MoE kernels and of the metadata refresh; swap-AB regression.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
timeout 400 python -m pytest tests/test_gpu_moe.py tests/test_gpu_metadata_update.py -q --maxfail=10 > gpurun_out/t_moe.log 2>&1; el "pytest moe + metadata rc=$?"; tail -15 gpurun_out/t_moe.log
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -q --maxfail=10 > gpurun_out/t_reg.log 2>&1; el "pytest gemm/model rc=$?"; tail -4 gpurun_out/t_reg.log
timeout 200 python - <<'PY'
import torch, math, sys
sys.path.insert(0, ".")
from xllm_b200 import ops
# MoE decode microbench: DeepSeek-V3-like routed experts at TP8 shard sizes? keep the reference's unquantised bf16 experts:
# Qwen3-30B-A3B shapes (H 2048, moe_intermediate 768, 128 experts, top-8), T = 1 and 16
dev = "cuda"
for T in (1, 16):
    H, I, E, k = 2048, 768, 128, 8
    x = torch.randn(T, H, device=dev, dtype=torch.bfloat16)
    fc1 = torch.randn(E, 2 * I, H, device=dev, dtype=torch.bfloat16) * 0.02
    fc2 = torch.randn(E, H, I, device=dev, dtype=torch.bfloat16) * 0.02
    logits = torch.randn(T, E, device=dev)
    w, ids = ops.moe_fused_topk(logits, k, True)
    out = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
    ws = torch.empty(T * k * (H + I) * 2, dtype=torch.uint8, device=dev)
    f = lambda: ops.cutlass_fused_moe(x, ids, w, fc1, fc2, output=out, workspace=ws)
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): f()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / 20
    nbytes = T * k * 3 * I * H * 2
    print(f"moe experts T={T} top-{k} of {E} (H {H}, I {I}): {us:.1f} us, {nbytes / us / 1e3:.0f} GB/s ({nbytes / us / 1e3 / 6482.4:.1%} of HBM peak; expert weights are L2-resident across iterations when T*k experts repeat)")
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(50): ops.moe_fused_topk(logits, k, True)
    t1.record(); torch.cuda.synchronize()
    print(f"moe_fused_topk T={T}: {t0.elapsed_time(t1) * 1e3 / 50:.2f} us")
PY
el "moe microbench rc=$?"
