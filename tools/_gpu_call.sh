#!/bin/bash
# One gpurun call: pair GEMM with CTA-scope barrier ops, prefill v2 with hoisted mask branch / try_wait hint / rescale threshold.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
B="--steps 20 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target"
timeout 600 python -m pytest tests/test_gpu_prefill_v2.py tests/test_gpu_prefill.py tests/test_gpu_gemm.py tests/test_gpu_q8.py tests/test_gpu_ffi.py -q --maxfail=8 > gpurun_out/t_tc.log 2>&1; el "pytest tc kernels rc=$?"; tail -5 gpurun_out/t_tc.log
XB_PREFILL_TAU=8 timeout 300 python -m pytest tests/test_gpu_prefill_v2.py -q --maxfail=8 -k "not bit_identical" > gpurun_out/t_v2_tau.log 2>&1; el "pytest v2 tau=8 rc=$?"; tail -5 gpurun_out/t_v2_tau.log
timeout 200 python tools/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1; el "gemm sweep rc=$?"; cat gpurun_out/gemm_sweep.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py $B > gpurun_out/bench_$tag.json 2>gpurun_out/bench_$tag.err; el "bench $tag rc=$?"; }
run v1 XB_PREFILL_V2=0
run v2 XB_PREFILL_V2=1
run v2tau8 XB_PREFILL_V2=1 XB_PREFILL_TAU=8
run w4pair XB_PREFILL_V2=1 XB_GEMM_CG=2
python - <<'PY'
import json
for f in ("v1", "v2", "v2tau8", "w4pair"):
    try:
        d = json.load(open(f"gpurun_out/bench_{f}.json"))
        print(f"{f:8s} tok/s {d['value']:7.1f} prefill lin TF {d['prefill']['linear_tflops']:.1f} attn TF {d['prefill']['attention_tflops_causal']:.1f} layer {d['prefill']['layer_tflops']:.1f}")
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/bench_{f}.err").read()[-600:])
PY
for t in prefill_v2; do
  XB_PREFILL_TAU=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"prefill_attention" -s 2 -c 1 -f -o /tmp/r02_$t python tools/profile_targets.py $t > gpurun_out/ncu_$t.log 2>&1; el "ncu $t rc=$?"
  ncu -i /tmp/r02_$t.ncu-rep --page raw --csv > gpurun_out/r02_${t}_tau8_raw.csv 2>/dev/null
  ncu -i /tmp/r02_$t.ncu-rep --page source --csv > gpurun_out/r02_${t}_tau8_source.csv 2>/dev/null
done
du -sh gpurun_out
