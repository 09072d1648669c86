#!/bin/bash
# One gpurun call: first run of the ping-pong prefill kernel; ncu of the slow W4 CTA-pair GEMM; exact-form tests.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
B="--steps 20 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target"
timeout 300 python -m pytest tests/test_gpu_prefill_v2.py -q --maxfail=8 > gpurun_out/t_v2.log 2>&1; el "pytest prefill v2 rc=$?"; tail -12 gpurun_out/t_v2.log
timeout 300 python -m pytest tests/test_gpu_w4_exact.py -q --maxfail=8 > gpurun_out/t_exact.log 2>&1; el "pytest exact rc=$?"; tail -4 gpurun_out/t_exact.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py $B > gpurun_out/bench_$tag.json 2>gpurun_out/bench_$tag.err; el "bench $tag rc=$?"; }
run base XB_PREFILL_V2=0
run v2 XB_PREFILL_V2=1
python - <<'PY'
import json
for f in ("base", "v2"):
    try:
        d = json.load(open(f"gpurun_out/bench_{f}.json"))
        print(f"{f:6s} tok/s {d['value']:7.1f} prefill lin TF {d['prefill']['linear_tflops']:.1f} attn TF {d['prefill']['attention_tflops_causal']:.1f} layer {d['prefill']['layer_tflops']:.1f}")
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/bench_{f}.err").read()[-600:])
PY
for t in gemm_w4_pair gemm_w4 prefill_v2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|prefill_attention" -s 2 -c 1 -f -o /tmp/r02_$t python tools/profile_targets.py $t > gpurun_out/ncu_$t.log 2>&1; el "ncu $t rc=$?"
  ncu -i /tmp/r02_$t.ncu-rep --page raw --csv > gpurun_out/r02_${t}_raw.csv 2>/dev/null
  ncu -i /tmp/r02_$t.ncu-rep --page source --csv > gpurun_out/r02_${t}_source.csv 2>/dev/null
done
du -sh gpurun_out
