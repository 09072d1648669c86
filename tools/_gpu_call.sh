#!/bin/bash
# One gpurun call: round-2 baseline evidence (bench line, launch list, ncu captures, the whole GPU suite, sweeps).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 900 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; el "bench default rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target > gpurun_out/bench_r02_short.json 2>/dev/null; el "bench short rc=$?"
XB_FUSE_GEMV=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target > gpurun_out/bench_r02_nofuse.json 2>/dev/null; el "bench nofuse rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"linear_|paged_decode|rms_norm|rope_and|embedding|argmax" -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target > gpurun_out/bench_under_ncu.log 2>&1; el "ncu list rc=$?"
for t in gate_up decode down qkv; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"linear_w4a16|paged_decode" -s 2 -c 1 -f -o /tmp/r02_$t python tools/profile_targets.py $t > gpurun_out/ncu_$t.log 2>&1; el "ncu $t rc=$?"
  ncu -i /tmp/r02_$t.ncu-rep --page raw --csv > gpurun_out/r02_${t}_raw.csv 2>/dev/null
  ncu -i /tmp/r02_$t.ncu-rep --page source --csv > gpurun_out/r02_${t}_source.csv 2>/dev/null
done
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -rs > gpurun_out/pytest_gpu.log 2>&1; el "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 200 python tools/decode_sweep.py b1 > gpurun_out/sweep_b1.log 2>&1; el "sweep rc=$?"
timeout 200 python tools/gemv_sweep.py 1 > gpurun_out/gemv.log 2>&1; el "gemv rc=$?"
du -sh gpurun_out
