#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
timeout 120 python -m pytest tests/test_gpu_model.py -q -k "fp8" > gpurun_out/r02k_pytest_fp8.log 2>&1; el "pytest rc=$?"; tail -4 gpurun_out/r02k_pytest_fp8.log
timeout 200 python bench.py > gpurun_out/r02k_bench.log 2>gpurun_out/r02k_bench.err; el "bench rc=$?"
grep '^{' gpurun_out/r02k_bench.log | tail -1 > gpurun_out/r02k_bench_line.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02k_bench_line.json"))
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches") if k in d})
print("scale_target", d.get("scale_target"))
PY
timeout 60 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02k_bench_reference.log 2>&1; el "reference rc=$?"; grep '^{' gpurun_out/r02k_bench_reference.log | tail -1 | cut -c1-300
