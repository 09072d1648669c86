#!/bin/bash
# Round-end refresh: tests of everything touched since the full-suite run, default bench line, reference arm, W4 pair capture.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
timeout 600 python -m pytest tests/test_gpu_moe.py tests/test_gpu_metadata_update.py tests/test_gpu_gemm.py tests/test_gpu_prefill_v2.py tests/test_gpu_model.py tests/test_gpu_model_prefill.py tests/test_gpu_ffi.py -q --maxfail=10 > gpurun_out/t_final2.log 2>&1; el "pytest rc=$?"; tail -4 gpurun_out/t_final2.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; el "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/bench_r02_final2.json 2> gpurun_out/bench_r02_final2.err; el "bench default rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_r02_reference2.json 2>/dev/null; el "bench reference rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:"gemm_tcgen05" -s 2 -c 1 -f -o /tmp/w4p2 python tools/profile_targets.py gemm_w4_pair > /dev/null 2>&1; ncu -i /tmp/w4p2.ncu-rep --page raw --csv > gpurun_out/r02h_gemm_w4_pair_raw.csv 2>/dev/null; el "ncu w4 pair rc=$?"
