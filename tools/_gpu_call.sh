mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target"
timeout 600 python -m pytest tests/test_gpu_fused_gemv.py tests/test_gpu_model.py tests/test_gpu_model_prefill.py tests/test_gpu_prefill.py tests/test_gpu_ffi.py tests/test_gpu_q8.py -q --maxfail=30 > gpurun_out/t5a.log 2>&1; echo "pytest-a rc=$?"; tail -3 gpurun_out/t5a.log
timeout 300 python bench.py $B > gpurun_out/bench5_default.json 2> gpurun_out/bench5_default.err; echo "bench default rc=$?"
XB_FUSE_GEMV=0 timeout 300 python bench.py $B > gpurun_out/bench5_nofuse.json 2>/dev/null; echo "bench nofuse rc=$?"
timeout 200 python tools/decode_sweep.py b1 > gpurun_out/sweep5_b1.log 2>&1; echo "sweep rc=$?"
timeout 200 python tools/gemv_sweep.py 1 > gpurun_out/gemv5.log 2>&1; echo "gemv rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"linear_|paged_decode|rms_norm|rope_and|embedding|argmax" -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-comparators --no-scale-target > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
for t in gate_up down qkv decode; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"linear_w4a16|paged_decode" -s 2 -c 1 -f -o /tmp/r02_$t python tools/profile_targets.py $t > gpurun_out/ncu_$t.log 2>&1; echo "ncu $t rc=$?"
  ncu -i /tmp/r02_$t.ncu-rep --page raw --csv > gpurun_out/r02_${t}_raw.csv 2>/dev/null
  ncu -i /tmp/r02_$t.ncu-rep --page source --csv > gpurun_out/r02_${t}_source.csv 2>/dev/null
done
timeout 1200 python -m pytest tests -m gpu -q --maxfail=60 > gpurun_out/t5.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t5.log
du -sh gpurun_out
