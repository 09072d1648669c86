#!/bin/bash
# One gpurun call: W4 pair GEMM with the deeper B ring, decode-attention merge changes, early PDL trigger, selective exact form.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
B="--steps 20 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target"
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_q8.py -q -x -k "cta_pair" > gpurun_out/t_pair.log 2>&1; el "pytest pair rc=$?"; tail -3 gpurun_out/t_pair.log
timeout 400 python -m pytest tests/test_gpu_decode.py tests/test_gpu_ffi.py tests/test_gpu_w4_exact.py -q --maxfail=10 > gpurun_out/t_dec.log 2>&1; el "pytest decode/ffi/exact rc=$?"; tail -5 gpurun_out/t_dec.log
timeout 200 python tools/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1; el "gemm sweep rc=$?"; cat gpurun_out/gemm_sweep.log
timeout 200 python tools/decode_sweep.py b1 > gpurun_out/sweep_b1.log 2>&1; el "sweep rc=$?"; head -3 gpurun_out/sweep_b1.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py $B > gpurun_out/bench_$tag.json 2>gpurun_out/bench_$tag.err; el "bench $tag rc=$?"; }
run base XB_W4_EXACT=0
run early XB_W4_EXACT=0 XB_PDL_EARLY=1
run exact2 XB_W4_EXACT=2
run exact2early XB_W4_EXACT=2 XB_PDL_EARLY=1
run w4pair XB_W4_EXACT=0 XB_GEMM_CG=2
python - <<'PY'
import json
for f in ("base", "early", "exact2", "exact2early", "w4pair"):
    try:
        d = json.load(open(f"gpurun_out/bench_{f}.json"))
        print(f"{f:12s} tok/s {d['value']:7.1f} ms {d['ms_per_step']:.4f} gate_up us {d['roofline']['launch_us']:.2f} frac {d['roofline']['frac']:.3f} "
              f"decode us {d['roofline']['paged_decode']['launch_us']:.2f} iso {d['roofline']['paged_decode']['launch_us_isolated']:.2f} "
              f"prefill lin TF {d['prefill']['linear_tflops']:.1f} attn {d['prefill']['attention_tflops_causal']:.1f}")
    except Exception as e:
        print(f, "failed", e)
PY
du -sh gpurun_out
