#!/bin/bash
# One gpurun call: where does the Llama-3-70B FP8 TP8 shard spend its step (one GPU, shard shapes, no exchange)?
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
for tp in 8 4 1; do timeout 300 python tools/shard_sim.py $tp 2>&1 | tail -1; el "sim tp$tp rc=$?"; done
XB_SIM_LAYERS=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_shard8_launches.csv python tools/shard_sim.py 8 > gpurun_out/shard8_ncu.log 2>&1; el "ncu shard8 rc=$?"
XB_SIM_LAYERS=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_shard1_launches.csv python tools/shard_sim.py 1 > gpurun_out/shard1_ncu.log 2>&1; el "ncu shard1 rc=$?"
python tools/ncu_summary.py list gpurun_out/r02_shard8_launches.csv gpurun_out/r02_shard8_launches.md "Llama-3-70B FP8 TP8 shard, B=32 ctx 8192, 4 layers" && cat gpurun_out/r02_shard8_launches.md
python tools/ncu_summary.py list gpurun_out/r02_shard1_launches.csv gpurun_out/r02_shard1_launches.md "Llama-3-70B FP8 unsharded, B=32 ctx 8192, 4 layers" && cat gpurun_out/r02_shard1_launches.md
