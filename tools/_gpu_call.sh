#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
timeout 200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_q8.py tests/test_gpu_model.py tests/test_gpu_model_prefill.py tests/test_gpu_linear.py -q --maxfail=8 > gpurun_out/r02j_pytest_subset.log 2>&1; el "pytest rc=$?"; tail -4 gpurun_out/r02j_pytest_subset.log
for tp in 8 1; do
  XB_FP8_NORM_QUANT=0 timeout 100 python tools/shard_sim.py $tp 2>&1 | tail -1
  XB_FP8_NORM_QUANT=1 timeout 100 python tools/shard_sim.py $tp 2>&1 | tail -1
done
el "sim done"
timeout 230 python bench.py > gpurun_out/r02j_bench.log 2>gpurun_out/r02j_bench.err; el "bench rc=$?"
grep '^{' gpurun_out/r02j_bench.log | tail -1 > gpurun_out/r02j_bench_line.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02j_bench_line.json"))
print({k: d[k] for k in ("value", "ms_per_step", "roofline", "e2e", "gpu_launches") if k in d})
print("scale_target", d.get("scale_target"))
print("comparators", d.get("comparators"))
PY
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; el "smoke"
