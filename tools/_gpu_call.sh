#!/bin/bash
# Final single-GPU evidence of the round: whole GPU suite, smoke, default bench line, reference arm, launch list, ncu captures.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -rs > gpurun_out/pytest_gpu_final.log 2>&1; el "pytest gpu rc=$?"; tail -6 gpurun_out/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; el "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; el "bench default rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_r02_reference.json 2>/dev/null; el "bench reference rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"linear_|paged_decode|rms_norm|rope_and|embedding|argmax" -c 800 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target > gpurun_out/bench_under_ncu.log 2>&1; el "ncu list rc=$?"
for t in gate_up decode down qkv gemm_w4_pair gemm_bf16_pair prefill_v2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"linear_w4a16|paged_decode|gemm_tcgen05|prefill_attention" -s 2 -c 1 -f -o /tmp/r02_$t python tools/profile_targets.py $t > gpurun_out/ncu_$t.log 2>&1; el "ncu $t rc=$?"
  ncu -i /tmp/r02_$t.ncu-rep --page raw --csv > gpurun_out/r02f_${t}_raw.csv 2>/dev/null
done
timeout 100 python tools/decode_sweep.py b1 2>&1 | head -3 > gpurun_out/sweep_b1_final.log; el "sweep rc=$?"
timeout 100 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_final.log 2>&1; el "gemm sweep rc=$?"
du -sh gpurun_out
