#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
timeout 200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_q8.py tests/test_gpu_model.py tests/test_gpu_model_prefill.py tests/test_gpu_linear.py -q --maxfail=8 > gpurun_out/r02j_pytest_subset.log 2>&1; el "pytest rc=$?"; tail -4 gpurun_out/r02j_pytest_subset.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; el "smoke"
timeout 230 python bench.py > gpurun_out/r02j_bench.log 2>gpurun_out/r02j_bench.err; el "bench rc=$?"
grep '^{' gpurun_out/r02j_bench.log | tail -1 > gpurun_out/r02j_bench_line.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02j_bench_line.json"))
print({k: d[k] for k in ("value", "ms_per_step", "roofline", "e2e", "gpu_launches") if k in d})
print("scale_target", d.get("scale_target"))
print("comparators", d.get("comparators"))
PY
timeout 100 python tools/fp8_splitk_probe.py 32 8 > gpurun_out/r02j_fp8_splitk_probe.txt 2>&1; el "probe rc=$?"; cat gpurun_out/r02j_fp8_splitk_probe.txt
