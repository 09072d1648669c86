#!/bin/bash
# One gpurun call: (1) first run of the CTA-pair GEMMs and the exact-dequant W4 GEMVs (under timeouts), (2) A/B numbers.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
B="--steps 20 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target"
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -x -k "cta_pair" > gpurun_out/t_pair.log 2>&1; el "pytest pair rc=$?"; tail -5 gpurun_out/t_pair.log
timeout 400 python -m pytest tests/test_gpu_w4_exact.py -q --maxfail=10 > gpurun_out/t_exact.log 2>&1; el "pytest exact rc=$?"; tail -8 gpurun_out/t_exact.log
timeout 200 python tools/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1; el "gemm sweep rc=$?"; cat gpurun_out/gemm_sweep.log
XB_SWEEP_FEW=1 XB_W4_EXACT=0 timeout 200 python tools/gemv_sweep.py 1 > gpurun_out/gemv_form0.log 2>&1; el "gemv form0 rc=$?"
XB_SWEEP_FEW=1 XB_W4_EXACT=1 timeout 200 python tools/gemv_sweep.py 1 > gpurun_out/gemv_form1.log 2>&1; el "gemv form1 rc=$?"; cat gpurun_out/gemv_form0.log gpurun_out/gemv_form1.log
XB_W4_EXACT=0 timeout 300 python bench.py $B > gpurun_out/bench_form0.json 2>gpurun_out/bench_form0.err; el "bench form0 rc=$?"
XB_W4_EXACT=1 timeout 300 python bench.py $B > gpurun_out/bench_form1.json 2>gpurun_out/bench_form1.err; el "bench form1 rc=$?"
python - <<'PY'
import json
for f in ("form0", "form1"):
    try:
        d = json.load(open(f"gpurun_out/bench_{f}.json"))
        print(f, "tok/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "gate_up us", round(d["roofline"]["launch_us"], 2), "frac", round(d["roofline"]["frac"], 3),
              "decode us", round(d["roofline"]["paged_decode"]["launch_us"], 2), "prefill lin TF", round(d["prefill"]["linear_tflops"], 1))
    except Exception as e:
        print(f, "failed", e)
PY
XB_GEMM_CG=2 timeout 300 python bench.py $B > gpurun_out/bench_pair.json 2>gpurun_out/bench_pair.err; el "bench pair rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_pair.json')); print('pair prefill', d['prefill'])"
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_linear.py tests/test_gpu_fused_gemv.py tests/test_gpu_model.py tests/test_gpu_q8.py -q --maxfail=20 > gpurun_out/t_regress.log 2>&1; el "pytest regress rc=$?"; tail -4 gpurun_out/t_regress.log
du -sh gpurun_out
