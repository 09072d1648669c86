"""Tensor-parallel parity check that needs no CPU oracle: the SAME logical weights are decoded once on a single GPU
(every rank computes it locally) and once through the TP-sharded runner; logits, greedy tokens and cross-rank
bit-identity are compared.  Used by tests/test_gpu_tp.py (torchrun, >= 2 GPUs) and by `bench.py --gpus N` (N > 1), which
prints the result as "tp_parity" before its timed region so that the driver's scaling run records TP correctness.

Sharding rules = the reference's (xllm/core/layers/common/qwen2_attention.cpp:47-65, linear.cpp:523-614,1405-1522); see
xllm_b200/parallel.py.
"""
import torch
import torch.distributed as dist

from . import parallel as P
from . import quant
from .qwen2 import Linear, Qwen2Config, Qwen2DecodeRunner, Qwen2Weights

BF16 = torch.bfloat16


def tiny_config(tp: int = 2) -> Qwen2Config:
    """small W4A16 stack whose head counts divide by tp (kv heads are replicated when tp > n_kv, as in the reference)."""
    return Qwen2Config(hidden_size=512, num_layers=3, n_heads=8, n_kv_heads=2, head_dim=64, intermediate_size=1024,
                       vocab_size=2048, block_size=16, quant="w4a16", group_size=64, max_position_embeddings=2048,
                       name="tiny")


def logical_weights(cfg: Qwen2Config, seed: int = 2026) -> dict:
    """logical (unsharded) weights on the CPU, identical on every rank: W4 linears as (q uint8 [N,K], s bf16 [N,K/g],
    z uint8 [N,K/g]) generated directly in quantised form, bf16 embedding / lm_head / norms."""
    g = torch.Generator().manual_seed(seed)
    H, I, gs = cfg.hidden_size, cfg.intermediate_size, cfg.group_size

    def lin(n, k, bias=False):
        q = torch.randint(0, 16, (n, k), generator=g, dtype=torch.uint8)
        s = ((torch.rand(n, k // gs, generator=g) * 0.5 + 0.75) * (0.05 * 3.0 / 7.5)).to(BF16)
        z = torch.randint(6, 10, (n, k // gs), generator=g, dtype=torch.uint8)
        b = (torch.randn(n, generator=g) * 0.05).to(BF16) if bias else None
        return dict(q=q, s=s, z=z, b=b)
    W = dict(embed=(torch.randn(cfg.vocab_size, H, generator=g) * 0.5).to(BF16),
             final_norm=(1 + 0.1 * torch.randn(H, generator=g)).to(BF16),
             lm_head=(torch.randn(cfg.vocab_size, H, generator=g) * 0.05).to(BF16), layers=[])
    for _ in range(cfg.num_layers):
        W["layers"].append(dict(input_norm=(1 + 0.1 * torch.randn(H, generator=g)).to(BF16),
                                post_norm=(1 + 0.1 * torch.randn(H, generator=g)).to(BF16),
                                qkv=lin(cfg.q_size + 2 * cfg.kv_size, H, cfg.qkv_bias), o=lin(H, cfg.q_size),
                                gate_up=lin(2 * I, H), down=lin(H, I)))
    return W


def shard_weights(cfg: Qwen2Config, W: dict, rank: int, tp: int, device, fuse_gate_up: bool = True, shard_embedding: bool = False):
    """this rank's Qwen2Weights (kernel layout, on `device`) from logical W4 weights; tp == 1 gives the full model.
    shard_embedding: the embedding table is split along the hidden dimension as the reference's WordEmbedding does
    (LOAD_SHARDED_WEIGHT(weight, 1), word_embedding_impl.cpp:60-64) instead of replicated."""
    hp = P.partition_heads(cfg.n_heads, cfg.n_kv_heads, rank, tp)
    w = Qwen2Weights(cfg)
    if shard_embedding and tp > 1:
        hs = cfg.hidden_size // tp
        w.embed = W["embed"][:, rank * hs:(rank + 1) * hs].contiguous().to(device)
    else:
        w.embed = W["embed"].to(device)
    w.final_norm = W["final_norm"].to(device)
    vs = cfg.vocab_size // tp
    w.lm_head = Linear(vs, cfg.hidden_size, "bf16")
    w.lm_head.weight = W["lm_head"][rank * vs:(rank + 1) * vs].contiguous().to(device)

    def mk(d, gate_up=False, qkv=False):
        n, k = d["q"].shape
        l = Linear(n, k, "w4a16", cfg.group_size)
        if qkv:
            qw, meta, b = quant.pack_w4_qkv_rope(d["q"], d["s"], d["z"], hp.num_heads, hp.num_kv_heads, cfg.head_dim,
                                                 cfg.group_size, d["b"])
            l.qkv_rope_packed = True
        elif gate_up and fuse_gate_up:
            qw, meta, b = quant.pack_w4_gate_up(d["q"], d["s"], d["z"], cfg.group_size, d["b"])
            l.gate_up_interleaved = True
        else:
            qw, meta = quant.pack_w4(d["q"], d["s"], d["z"], cfg.group_size)
            b = d["b"]
        l.qweight, l.meta = qw.to(device), meta.to(device)
        l.bias = b.to(device) if b is not None else None
        return l
    gs = cfg.group_size
    for L in W["layers"]:
        qkv = P.shard_linear("w4", L["qkv"], P.shard_qkv_rows(cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, rank, tp), None, gs, rank)
        o = P.shard_linear("w4", L["o"], None, P.shard_cols(cfg.q_size, rank, tp), gs, rank)
        gu = P.shard_linear("w4", L["gate_up"], P.shard_gate_up_rows(cfg.intermediate_size, rank, tp), None, gs, rank)
        dn = P.shard_linear("w4", L["down"], None, P.shard_cols(cfg.intermediate_size, rank, tp), gs, rank)
        w.layers.append(dict(input_norm=L["input_norm"].to(device), post_norm=L["post_norm"].to(device), qkv=mk(qkv, qkv=True),
                             o=mk(o), gate_up=mk(gu, True), down=mk(dn)))
    return w, hp


def decode_case(cfg: Qwen2Config, kv_lens, seed: int = 7):
    """prefilled caches (full kv heads, CPU) + the integer step inputs of one decode step."""
    g = torch.Generator().manual_seed(seed)
    bs = cfg.block_size
    npg = [(n + bs - 1) // bs for n in kv_lens]
    nblocks = sum(npg) + 3
    perm = (torch.randperm(nblocks - 1, generator=g) + 1)[:sum(npg)].tolist()
    indptr = [0]
    for n in npg:
        indptr.append(indptr[-1] + n)
    B = len(kv_lens)
    slots = [perm[indptr[b] + (kv_lens[b] - 1) // bs] * bs + (kv_lens[b] - 1) % bs for b in range(B)]
    kcs = [torch.randn(nblocks, bs, cfg.n_kv_heads, cfg.head_dim, generator=g).to(BF16) for _ in range(cfg.num_layers)]
    vcs = [torch.randn(nblocks, bs, cfg.n_kv_heads, cfg.head_dim, generator=g).to(BF16) for _ in range(cfg.num_layers)]
    meta = dict(tokens=torch.randint(0, cfg.vocab_size, (B,), generator=g).tolist(), positions=[n - 1 for n in kv_lens],
                slots=slots, indptr=indptr, indices=perm, last=[(n - 1) % bs + 1 for n in kv_lens], nblocks=nblocks)
    return kcs, vcs, meta


def _run(cfg, w, hp, kcs, vcs, meta, B, max_ctx, dev, pg, exchange, use_graph):
    run = Qwen2DecodeRunner(cfg, w, B, max_ctx, device=dev, num_blocks=meta["nblocks"], pg=pg, exchange=exchange)
    sl = slice(hp.kv_head0, hp.kv_head0 + hp.num_kv_heads)

    def fill():
        for li in range(cfg.num_layers):
            run.k_caches[li].copy_(kcs[li][:, :, sl])
            run.v_caches[li].copy_(vcs[li][:, :, sl])
    fill()
    run.set_inputs_host(meta["tokens"], meta["positions"], meta["slots"], meta["indptr"], meta["indices"], meta["last"])
    run.step()
    if use_graph:
        fill()
        run.capture()                   # NCCL / symmetric-memory kernels inside the CUDA graph
        run.step()
    return run.h_next.clone(), run.logits.clone(), run


def tp_parity(pg: P.ProcessGroup, device, exchange: str = "peer", kv_lens=(37, 300, 1), use_graph: bool = True) -> dict:
    """-> {"rel_l2", "tokens_equal", "ranks_bit_identical", "exchange", "tp"}; collective over `pg` (all ranks call it)."""
    tp, rank = pg.world_size, pg.rank
    cfg = tiny_config(tp)
    W = logical_weights(cfg)
    kv_lens = list(kv_lens)
    B = len(kv_lens)
    kcs, vcs, meta = decode_case(cfg, kv_lens)
    w1, hp1 = shard_weights(cfg, W, 0, 1, device)
    ref_next, ref_logits, _ = _run(cfg, w1, hp1, kcs, vcs, meta, B, max(kv_lens), device, None, "nccl", use_graph)
    w, hp = shard_weights(cfg, W, rank, tp, device)
    nxt, logits, run = _run(cfg, w, hp, kcs, vcs, meta, B, max(kv_lens), device, pg, exchange, use_graph)
    rel = ((logits.float() - ref_logits.float()).norm() / ref_logits.float().norm()).item()
    gathered = [torch.empty_like(logits) for _ in range(tp)]
    dist.all_gather(gathered, logits, group=pg.group)
    same = all(torch.equal(gathered[0], t) for t in gathered[1:])
    flags = torch.tensor([float(rel), float(torch.equal(nxt[:B], ref_next[:B])), float(same)], device=device, dtype=torch.float64)
    mx, mn = flags.clone(), flags.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=pg.group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=pg.group)
    return {"rel_l2": float(mx[0]), "tokens_equal": bool(mn[1] > 0.5), "ranks_bit_identical": bool(mn[2] > 0.5),
            "exchange": run.exchange_mode, "tp": tp, "config": "3-layer W4A16 stack (H=512, 8/2 heads), batch 3, graph replay"}


def tp_prefill_parity(pg: P.ProcessGroup, device, lens=(37, 130, 5), shard_embedding: bool = True) -> dict:
    """prompt prefill (ragged, first chunk) through the TP-sharded Qwen2PrefillRunner vs the single-GPU runner on the same
    logical weights: last-token logits, greedy tokens, the KV rows this rank owns, cross-rank bit-identity.  Collective."""
    from .qwen2_prefill import Qwen2PrefillRunner
    tp, rank = pg.world_size, pg.rank
    cfg = tiny_config(tp)
    W = logical_weights(cfg)
    lens = list(lens)
    B, T = len(lens), sum(lens)
    g = torch.Generator().manual_seed(11)
    bs = cfg.block_size
    npg = [(n + bs - 1) // bs for n in lens]
    nblocks = sum(npg) + 2
    perm = (torch.randperm(nblocks - 1, generator=g) + 1)[:sum(npg)].tolist()
    tokens = torch.randint(0, cfg.vocab_size, (T,), generator=g, dtype=torch.int32)
    positions = torch.cat([torch.arange(n) for n in lens]).to(torch.int64)
    slots, off = [], 0
    for b, n in enumerate(lens):
        for i in range(n):
            slots.append(perm[off + i // bs] * bs + i % bs)
        off += npg[b]
    slots = torch.tensor(slots, dtype=torch.int32)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)

    def run(w, hp, group):
        r = Qwen2DecodeRunner(cfg, w, B, max(lens) + 1, device=device, num_blocks=nblocks, pg=group, exchange="nccl")
        pr = Qwen2PrefillRunner.from_decode_runner(r)
        logits, toks = pr.forward(tokens.to(device), positions.to(device), slots.to(device), cu.to(device), max_qo_len=max(lens))
        torch.cuda.synchronize()
        return logits, toks, r
    w1, hp1 = shard_weights(cfg, W, 0, 1, device)
    ref_logits, ref_toks, r1 = run(w1, hp1, None)
    w, hp = shard_weights(cfg, W, rank, tp, device, shard_embedding=shard_embedding)
    logits, toks, r = run(w, hp, pg)
    rel = ((logits.float() - ref_logits.float()).norm() / ref_logits.float().norm()).item()
    sl = slice(hp.kv_head0, hp.kv_head0 + hp.num_kv_heads)
    kv_rel = 0.0
    for li in range(cfg.num_layers):
        a, b_ = r.k_caches[li].float(), r1.k_caches[li][:, :, sl].float()
        kv_rel = max(kv_rel, ((a - b_).norm() / b_.norm().clamp_min(1e-9)).item())
    gathered = [torch.empty_like(logits) for _ in range(tp)]
    dist.all_gather(gathered, logits, group=pg.group)
    same = all(torch.equal(gathered[0], t) for t in gathered[1:])
    flags = torch.tensor([float(rel), float(kv_rel), float(torch.equal(toks, ref_toks)), float(same)], device=device, dtype=torch.float64)
    mx, mn = flags.clone(), flags.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=pg.group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=pg.group)
    return {"rel_l2": float(mx[0]), "kv_rel_l2": float(mx[1]), "tokens_equal": bool(mn[2] > 0.5), "ranks_bit_identical": bool(mn[3] > 0.5),
            "tp": tp, "sharded_embedding": bool(shard_embedding and tp > 1)}
