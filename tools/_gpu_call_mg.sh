#!/bin/bash
# multi-GPU validation call: TP parity tests (decode + prefill + sharded embedding) under torchrun
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  -m pytest tests/test_gpu_tp.py -q -m gpu -x > gpurun_out/r02h_tp_tests.log 2>&1
echo "tp tests rc=$?"
tail -15 gpurun_out/r02h_tp_tests.log
