#!/bin/bash
# Multi-GPU validation (gpurun --gpus 4): TP parity test, bench at N=2, N=4 as two TP2 sub-groups (the N=8 layout in small), N=4 TP4.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--steps 20 --warmup 3 --no-cpu-baseline --no-comparators"
export XB_BENCH_WATCHDOG=240
timeout 300 $TR --nproc-per-node 2 --master-port 29501 -m pytest tests/test_gpu_tp.py -q -m gpu > gpurun_out/mg_tp_test.log 2>&1; el "tp test rc=$?"; tail -3 gpurun_out/mg_tp_test.log
timeout 400 $TR --nproc-per-node 2 --master-port 29502 bench.py --gpus 2 $B > gpurun_out/mg_bench_n2.json 2> gpurun_out/mg_bench_n2.err; el "bench n2 rc=$?"
timeout 400 $TR --nproc-per-node 4 --master-port 29503 bench.py --gpus 4 --tp 2 $B --no-scale-target > gpurun_out/mg_bench_n4_tp2dp2.json 2> gpurun_out/mg_bench_n4_tp2dp2.err; el "bench n4 tp2xdp2 rc=$?"
timeout 500 $TR --nproc-per-node 4 --master-port 29504 bench.py --gpus 4 $B > gpurun_out/mg_bench_n4.json 2> gpurun_out/mg_bench_n4.err; el "bench n4 rc=$?"
python - <<'PY'
import json
for f in ("n2", "n4_tp2dp2", "n4"):
    try:
        d = json.load(open(f"gpurun_out/mg_bench_{f}.json"))
        st = d.get("scale_target") or {}
        print(f"{f:10s} tok/s {d['value']:7.1f} ms {d['ms_per_step']:.4f} par {d['config']['parallelism']} exch {d['config']['exchange']} tp_parity {d.get('tp_parity')}")
        print("           scale_target", {k: st.get(k) for k in ("tp", "tokens_per_s", "ms_per_step", "hbm_frac_per_gpu", "exchange", "error")})
    except Exception as e:
        print(f, "failed", e)
        try:
            print(open(f"gpurun_out/mg_bench_{f}.err").read()[-1500:])
        except Exception:
            pass
PY
