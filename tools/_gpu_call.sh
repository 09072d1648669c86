#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
B="--steps 20 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target"
XB_PREFILL_QK_FIRST=1 XB_PREFILL_SPIN=1 timeout 300 python -m pytest tests/test_gpu_prefill_v2.py tests/test_gpu_moe.py -q --maxfail=8 > gpurun_out/t_v2.log 2>&1; el "pytest v2 (qk first, spin) + moe rc=$?"; tail -4 gpurun_out/t_v2.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py $B > gpurun_out/bench_$tag.json 2>gpurun_out/bench_$tag.err; el "bench $tag rc=$?"; }
run base XB_PREFILL_QK_FIRST=0 XB_PREFILL_SPIN=0
run qkfirst XB_PREFILL_QK_FIRST=1 XB_PREFILL_SPIN=0
run spin XB_PREFILL_QK_FIRST=0 XB_PREFILL_SPIN=1
run both XB_PREFILL_QK_FIRST=1 XB_PREFILL_SPIN=1
run bothtau XB_PREFILL_QK_FIRST=1 XB_PREFILL_SPIN=1 XB_PREFILL_TAU=8
python - <<'PY'
import json
for f in ("base", "qkfirst", "spin", "both", "bothtau"):
    try:
        d = json.load(open(f"gpurun_out/bench_{f}.json"))
        print(f"{f:8s} tok/s {d['value']:7.1f} prefill lin TF {d['prefill']['linear_tflops']:.1f} attn TF {d['prefill']['attention_tflops_causal']:.1f}")
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/bench_{f}.err").read()[-400:])
PY
