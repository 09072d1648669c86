mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target"
timeout 400 python -m pytest tests/test_gpu_fused_gemv.py tests/test_gpu_q8.py tests/test_gpu_decode.py tests/test_gpu_linear.py -q --maxfail=30 > gpurun_out/t3a.log 2>&1; echo "pytest-a rc=$?"; tail -3 gpurun_out/t3a.log
timeout 200 python tools/decode_sweep.py b1 > gpurun_out/sweep3_b1.log 2>&1; echo "sweep rc=$?"
XB_W4_EXACT=1 timeout 200 python tools/gemv_sweep.py 1 > gpurun_out/gemv3_exact.log 2>&1; echo "gemv exact rc=$?"
XB_W4_EXACT=0 timeout 200 python tools/gemv_sweep.py 1 > gpurun_out/gemv3_bf16.log 2>&1; echo "gemv bf16 rc=$?"
timeout 300 python tools/gemv_sweep.py q8 > gpurun_out/gemv3_q8.log 2>&1; echo "gemv q8 rc=$?"
timeout 300 python bench.py $B > gpurun_out/bench3_default.json 2> gpurun_out/bench3_default.err; echo "bench default rc=$?"
XB_FUSE_GEMV=0 timeout 300 python bench.py $B > gpurun_out/bench3_nofuse.json 2>/dev/null; echo "bench nofuse rc=$?"
XB_SMEM_CARVEOUT=0 timeout 300 python bench.py $B > gpurun_out/bench3_nocarve.json 2>/dev/null; echo "bench nocarve rc=$?"
XB_FUSE_GEMV=0 XB_SMEM_CARVEOUT=0 timeout 300 python bench.py $B > gpurun_out/bench3_nofuse_nocarve.json 2>/dev/null; echo "rc=$?"
XB_W4_EXACT=0 timeout 300 python bench.py $B > gpurun_out/bench3_noexact.json 2>/dev/null; echo "rc=$?"
XB_W4_EXACT=0 XB_FUSE_GEMV=0 timeout 300 python bench.py $B > gpurun_out/bench3_noexact_nofuse.json 2>/dev/null; echo "rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
timeout 1200 python -m pytest tests -m gpu -q --maxfail=60 > gpurun_out/t3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t3.log
