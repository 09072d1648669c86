#!/bin/bash
# Multi-GPU validation 2 (gpurun --gpus 4): the N=8 sequence in small - TP sub-groups for the headline, THEN the scale target on the
# world group (a second symmetric-memory rendezvous on a different group) - and N=4 TP4 with the swap-AB FP8 GEMMs.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
el() { echo "[+$(( $(date +%s) - t0 )) s] $*"; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--steps 20 --warmup 3 --no-cpu-baseline --no-comparators"
export XB_BENCH_WATCHDOG=300 XB_SCALE_TARGET_TIMEOUT=200
timeout 500 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --tp 2 $B > gpurun_out/mg2_bench_n4_tp2dp2.json 2> gpurun_out/mg2_bench_n4_tp2dp2.err; el "bench n4 tp2xdp2 + scale target rc=$?"
timeout 500 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 $B > gpurun_out/mg2_bench_n4.json 2> gpurun_out/mg2_bench_n4.err; el "bench n4 rc=$?"
timeout 400 $TR --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 $B > gpurun_out/mg2_bench_n2.json 2> gpurun_out/mg2_bench_n2.err; el "bench n2 rc=$?"
python - <<'PY'
import json
for f in ("n4_tp2dp2", "n4", "n2"):
    try:
        txt = open(f"gpurun_out/mg2_bench_{f}.json").read()
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        st = d.get("scale_target") or {}
        print(f"{f:10s} tok/s {d['value']:7.1f} ms {d['ms_per_step']:.4f} par {d['config']['parallelism']} exch {d['config']['exchange']} tp_parity ok={d['tp_parity']['tokens_equal'] and d['tp_parity']['ranks_bit_identical']}")
        print("           scale_target", {k: st.get(k) for k in ("tp", "tokens_per_s", "ms_per_step", "hbm_frac_per_gpu", "exchange", "error")})
    except Exception as e:
        print(f, "failed", e)
        try:
            print("\n".join(l for l in open(f"gpurun_out/mg2_bench_{f}.err").read().splitlines() if "FutureWarning" not in l and "symm_mem.enable" not in l)[-1500:])
        except Exception:
            pass
PY
