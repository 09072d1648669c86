"""Multi-GPU parity (run with `gpurun --gpus 2 -- torchrun --nproc-per-node 2 -m pytest tests/test_gpu_tp.py -m gpu`):
a TP=2 decode step (NCCL baseline and the NVLink one-shot fused exchange) against the single-GPU runner."""
import os

import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu


def _init():
    if "RANK" not in os.environ or int(os.environ.get("WORLD_SIZE", "1")) < 2:
        pytest.skip("needs torchrun with >= 2 ranks")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_tp_decode_step_matches_single_gpu(exchange, built_lib):
    _init()
    from xllm_b200 import parallel as P
    from xllm_b200.tp_check import tp_parity
    dev = f"cuda:{torch.cuda.current_device()}"
    r = tp_parity(P.ProcessGroup(), dev, exchange)
    assert r["rel_l2"] <= 2e-2, r
    assert r["tokens_equal"], "greedy tokens differ between TP and single GPU"
    assert r["ranks_bit_identical"], "ranks disagree (fixed rank order in the one-shot exchange)"
    assert exchange not in r["exchange"] or r["exchange"] == exchange, r
