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


@pytest.mark.parametrize("shard_embedding", [True, False])
def test_tp_prefill_matches_single_gpu(shard_embedding, built_lib):
    """TP prompt prefill (row-parallel all-reduces at T > 1, gathered lm_head, embedding sharded along the hidden dimension as
    word_embedding_impl.cpp:48-56) against the single-GPU prefill runner."""
    _init()
    from xllm_b200 import parallel as P
    from xllm_b200.tp_check import tp_prefill_parity
    dev = f"cuda:{torch.cuda.current_device()}"
    r = tp_prefill_parity(P.ProcessGroup(), dev, shard_embedding=shard_embedding)
    assert r["rel_l2"] <= 2e-2 and r["kv_rel_l2"] <= 2e-2, r
    assert r["tokens_equal"] and r["ranks_bit_identical"], r


def test_tp_decode_with_sharded_embedding(built_lib):
    """decode step with the hidden-sharded embedding table (lookup + all-gather inside the CUDA graph)."""
    _init()
    from xllm_b200 import parallel as P
    from xllm_b200 import tp_check as TC
    dev = f"cuda:{torch.cuda.current_device()}"
    pg = P.ProcessGroup()
    cfg = TC.tiny_config(pg.world_size)
    W = TC.logical_weights(cfg)
    kcs, vcs, meta = TC.decode_case(cfg, [37, 300, 1])
    w1, hp1 = TC.shard_weights(cfg, W, 0, 1, dev)
    ref_next, ref_logits, _ = TC._run(cfg, w1, hp1, kcs, vcs, meta, 3, 300, dev, None, "nccl", True)
    w, hp = TC.shard_weights(cfg, W, pg.rank, pg.world_size, dev, shard_embedding=True)
    nxt, logits, run = TC._run(cfg, w, hp, kcs, vcs, meta, 3, 300, dev, pg, "peer", True)
    rel = ((logits.float() - ref_logits.float()).norm() / ref_logits.float().norm()).item()
    assert rel <= 2e-2 and torch.equal(nxt[:3], ref_next[:3]), rel
