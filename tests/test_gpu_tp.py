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


def _cfg():
    from xllm_b200.qwen2 import Qwen2Config
    return Qwen2Config(hidden_size=512, num_layers=3, n_heads=8, n_kv_heads=2, head_dim=64, intermediate_size=1024,
                       vocab_size=2048, block_size=16, quant="w4a16", group_size=64, max_position_embeddings=2048, name="tiny")


def _shard_weights(cfg, W, rank, tp, device):
    from xllm_b200 import parallel as P
    from xllm_b200 import quant
    from xllm_b200.qwen2 import Linear, Qwen2Weights
    hp = P.partition_heads(cfg.n_heads, cfg.n_kv_heads, rank, tp)
    w = Qwen2Weights(cfg)
    w.embed = W["embed"].to(device)
    w.final_norm = W["final_norm"].to(device)
    vs = cfg.vocab_size // tp
    w.lm_head = Linear(vs, cfg.hidden_size, "bf16")
    w.lm_head.weight = W["lm_head"][rank * vs:(rank + 1) * vs].contiguous().to(device)

    def mk(d):
        n, k = d["q"].shape
        l = Linear(n, k, "w4a16", cfg.group_size)
        qw, meta = quant.pack_w4(d["q"], d["s"], d["z"], cfg.group_size)
        l.qweight, l.meta = qw.to(device), meta.to(device)
        l.bias = d["b"].to(device) if d["b"] is not None else None
        return l
    for L in W["layers"]:
        qkv = P.shard_linear("w4", L["qkv"], P.shard_qkv_rows(cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, rank, tp), None, cfg.group_size)
        o = P.shard_linear("w4", L["o"], None, P.shard_cols(cfg.q_size, rank, tp), cfg.group_size)
        gu = P.shard_linear("w4", L["gate_up"], P.shard_gate_up_rows(cfg.intermediate_size, rank, tp), None, cfg.group_size)
        dn = P.shard_linear("w4", L["down"], None, P.shard_cols(cfg.intermediate_size, rank, tp), cfg.group_size)
        w.layers.append(dict(input_norm=L["input_norm"].to(device), post_norm=L["post_norm"].to(device), qkv=mk(qkv), o=mk(o),
                             gate_up=mk(gu), down=mk(dn)))
    return w, hp


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_tp_decode_step_matches_single_gpu(exchange, built_lib):
    _init()
    from tests.model_parity import build_case, upload
    from tests.util import assert_close_bf16
    from xllm_b200 import parallel as P
    from xllm_b200.qwen2 import Qwen2DecodeRunner
    cfg = _cfg()
    rank, tp = dist.get_rank(), dist.get_world_size()
    dev = f"cuda:{torch.cuda.current_device()}"
    kv_lens = [37, 300, 1]
    W, kcs, vcs, meta = build_case(cfg, 3, kv_lens)
    # single-GPU result (every rank computes it locally)
    ref = Qwen2DecodeRunner(cfg, upload(cfg, W, dev), 3, max(kv_lens), device=dev, num_blocks=meta["nblocks"])
    for li in range(cfg.num_layers):
        ref.k_caches[li].copy_(kcs[li]); ref.v_caches[li].copy_(vcs[li])
    ref.set_inputs_host(meta["tokens"], meta["positions"], meta["slots"], meta["indptr"], meta["indices"], meta["last"])
    ref_next = ref.step().clone()
    ref_logits = ref.logits.clone()
    # TP result
    pg = P.ProcessGroup()
    w, hp = _shard_weights(cfg, W, rank, tp, dev)
    run = Qwen2DecodeRunner(cfg, w, 3, max(kv_lens), device=dev, num_blocks=meta["nblocks"], pg=pg, exchange=exchange)
    sl = slice(hp.kv_head0, hp.kv_head0 + hp.num_kv_heads)
    for li in range(cfg.num_layers):
        run.k_caches[li].copy_(kcs[li][:, :, sl]); run.v_caches[li].copy_(vcs[li][:, :, sl])
    run.set_inputs_host(meta["tokens"], meta["positions"], meta["slots"], meta["indptr"], meta["indices"], meta["last"])
    run.step()                       # eager
    for li in range(cfg.num_layers):
        run.k_caches[li].copy_(kcs[li][:, :, sl]); run.v_caches[li].copy_(vcs[li][:, :, sl])
    run.capture()                    # then as a CUDA graph (NCCL / symmetric-memory kernels inside the graph)
    nxt = run.step().clone()
    assert_close_bf16(run.logits, ref_logits, ulps=1e9, rel_l2=2e-2, what=f"TP{tp} logits ({exchange})")
    assert torch.equal(nxt[:3], ref_next[:3]), "greedy tokens differ between TP and single GPU"
    # all ranks must agree bit-for-bit (fixed rank order in the one-shot exchange)
    gathered = [torch.empty_like(run.logits) for _ in range(tp)]
    dist.all_gather(gathered, run.logits)
    assert all(torch.equal(gathered[0], g) for g in gathered[1:]), "ranks disagree"
