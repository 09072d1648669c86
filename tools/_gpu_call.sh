#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
B="--steps 30 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target"
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_linear.py -q --maxfail=8 -k "mlp_norm or checkpoint_layout or matches_oracle" > gpurun_out/t_misc.log 2>&1; el "pytest rc=$?"; tail -3 gpurun_out/t_misc.log
for i in 1 2; do
XB_FUSE_MLP_NORM=0 timeout 200 python bench.py $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base    ', round(d['value'],1), round(d['ms_per_step'],4), d['config']['launches_per_step'])"
XB_FUSE_MLP_NORM=1 timeout 200 python bench.py $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mlp_norm', round(d['value'],1), round(d['ms_per_step'],4), d['config']['launches_per_step'])"
done
el "bench A/B done"
