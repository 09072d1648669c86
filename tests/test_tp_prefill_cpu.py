"""CPU test (gloo, world_size 2) of the TENSOR-PARALLEL prefill composition (xllm_b200/qwen2_prefill.py under a process group):
sharded heads / intermediate columns, the all-reduce after every row-parallel linear at T > 1 (linear.cpp:1518-1520), the gathered
column-parallel lm_head, and the embedding table split along the hidden dimension + all-gather (word_embedding_impl.cpp:48-64) -
with the library ops replaced by the oracle adapter of tests/test_prefill_composition_cpu.py.  Checked against the single-rank
oracle composition; layer 0's K/V (computed before any exchange) must be bit-identical to the oracle's rows for this rank's heads."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import batch as OB
from tests import model_parity as MP
from tests.test_parallel_cpu import _free_port
from tests.test_prefill_composition_cpu import OracleOps, _fresh_caches
from xllm_b200 import parallel as P
from xllm_b200 import qwen2 as Q2
from xllm_b200.qwen2 import Linear, Qwen2Config, Qwen2Weights
from xllm_b200.qwen2_prefill import Qwen2PrefillRunner

BF16 = torch.bfloat16
LENS = [5, 9, 1]
BLOCKS = [[3, 1], [6, 2, 5], [4]]
NBLOCKS = 8


def _cfg():
    return Qwen2Config(hidden_size=64, num_layers=2, n_heads=4, n_kv_heads=2, head_dim=16, intermediate_size=96,
                       vocab_size=128, max_position_embeddings=64, block_size=4, quant="bf16", name="tiny-tp")


def _case(cfg):
    W, _, _, _ = MP.build_case(cfg, 1, [1], seed=7)
    g = torch.Generator().manual_seed(3)
    toks = [t for n in LENS for t in torch.randint(0, cfg.vocab_size, (n,), generator=g).tolist()]
    meta = OB.build_paged_meta([OB.SeqState(b, 0, n) for b, n in zip(BLOCKS, LENS)], cfg.block_size)
    return W, toks, meta


def _shard(cfg, W, rank, tp, shard_embedding):
    """this rank's bf16 weights: the same row / column selections the W4 path uses (parallel.shard_* index helpers)"""
    H, I, D = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    hp = P.partition_heads(cfg.n_heads, cfg.n_kv_heads, rank, tp)
    w = Qwen2Weights(cfg)
    hs = H // tp
    w.embed = W["embed"][:, rank * hs:(rank + 1) * hs].contiguous() if shard_embedding else W["embed"]
    w.final_norm = W["final_norm"]
    vs = cfg.vocab_size // tp
    w.lm_head = Linear(vs, H, "bf16")
    w.lm_head.weight = W["lm_head"][rank * vs:(rank + 1) * vs].contiguous()

    def lin(d, rows=None, cols=None):
        wt = d["w"]
        if rows is not None:
            wt = wt[rows]
        if cols is not None:
            wt = wt[:, cols]
        l = Linear(wt.size(0), wt.size(1), "bf16")
        l.weight = wt.contiguous()
        l.bias = d["b"][rows].contiguous() if (d["b"] is not None and rows is not None) else None
        return l
    for L in W["layers"]:
        w.layers.append(dict(input_norm=L["input_norm"], post_norm=L["post_norm"],
                             qkv=lin(L["qkv"], rows=P.shard_qkv_rows(cfg.n_heads, cfg.n_kv_heads, D, rank, tp)),
                             o=lin(L["o"], cols=P.shard_cols(cfg.n_heads * D, rank, tp)),
                             gate_up=lin(L["gate_up"], rows=P.shard_gate_up_rows(I, rank, tp)),
                             down=lin(L["down"], cols=P.shard_cols(I, rank, tp))))
    return w, hp


def _worker(rank, world, port, shard_embedding, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Q2.ops = OracleOps                                   # Linear.forward and the runner resolve ops through qwen2
    cfg = _cfg()
    W, toks, meta = _case(cfg)
    w, hp = _shard(cfg, W, rank, world, shard_embedding)
    mk = lambda: [torch.zeros(NBLOCKS, cfg.block_size, hp.num_kv_heads, cfg.head_dim, dtype=BF16) for _ in range(cfg.num_layers)]
    kcs, vcs = mk(), mk()
    r = Qwen2PrefillRunner(cfg, w, kcs, vcs, Q2.make_cos_sin_cache(cfg, "cpu"), device="cpu", pg=P.ProcessGroup())
    assert (r.nh, r.nkv, r.inter) == (2, 1, 48)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32)
    logits, tokens = r.forward(i32(toks), torch.tensor(meta.positions, dtype=torch.int64), i32(meta.new_cache_slots),
                               i32(meta.q_cu_seq_lens), i32(meta.kv_cu_seq_lens))
    # the same prompts in two chunks ([first 3 | rest], sequence 2 finishes in chunk 1) through the TP runner: chunked prefill over
    # the paged cache must reproduce the one-shot TP result bit for bit (rows are independent in every exchange)
    cut = [3, 3, 1]
    bs = cfg.block_size
    kc2, vc2 = mk(), mk()
    r2 = Qwen2PrefillRunner(cfg, w, kc2, vc2, Q2.make_cos_sin_cache(cfg, "cpu"), device="cpu", pg=P.ProcessGroup())
    seqs = []
    off = 0
    for n in LENS:
        seqs.append(toks[off:off + n])
        off += n
    m1 = OB.build_paged_meta([OB.SeqState(b[: (c + bs - 1) // bs], 0, c) for b, c in zip(BLOCKS, cut)], bs)
    first = [t for sq, c in zip(seqs, cut) for t in sq[:c]]
    r2.forward(i32(first), torch.tensor(m1.positions, dtype=torch.int64), i32(m1.new_cache_slots), i32(m1.q_cu_seq_lens),
               i32(m1.kv_cu_seq_lens))
    rest = [(b, c, n) for b, c, n in zip(BLOCKS, cut, LENS) if n > c]
    m2 = OB.build_paged_meta([OB.SeqState(b, c, n) for b, c, n in rest], bs)
    second = [t for sq, c, n in zip(seqs, cut, LENS) if n > c for t in sq[c:]]
    logits2, tokens2 = r2.forward(i32(second), torch.tensor(m2.positions, dtype=torch.int64), i32(m2.new_cache_slots),
                                  i32(m2.q_cu_seq_lens), None, i32(m2.paged_kv_indptr), i32(m2.paged_kv_indices),
                                  i32(m2.paged_kv_last_page_len), chunked=True)
    chunk_ok = bool(torch.equal(logits2, logits[:2]) and torch.equal(tokens2, tokens[:2]) and
                    all(torch.equal(a, b) for a, b in zip(kc2 + vc2, kcs + vcs)))
    out_q.put((rank, logits.float().numpy(), tokens.numpy(), [k.float().numpy() for k in kcs], [v.float().numpy() for v in vcs], chunk_ok))
    dist.barrier()
    dist.destroy_process_group()


def _run(shard_embedding):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shard_embedding, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(world):
        item = q.get(timeout=180)
        got[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return got


def test_tp2_prefill_composition_matches_single_rank_oracle():
    cfg = _cfg()
    W, toks, meta = _case(cfg)
    kc_o, vc_o = _fresh_caches(cfg, NBLOCKS)
    ref = MP.oracle_prefill(cfg, W, kc_o, vc_o, toks, meta, chunked=False)
    for shard_embedding in (True, False):
        got = _run(shard_embedding)
        l0, l1 = torch.from_numpy(got[0][0]), torch.from_numpy(got[1][0])
        assert torch.equal(l0, l1), "ranks disagree on the gathered logits"
        assert l0.shape == ref.shape
        rel = ((l0 - ref.float()).norm() / ref.float().norm()).item()
        # per-rank partial sums are rounded to bf16 before the exchange (the reference's NCCL all-reduce does the same)
        assert rel <= 2e-2, f"TP2 prefill logits rel-L2 {rel:.3e}"
        assert torch.equal(torch.from_numpy(got[0][1]).long(), ref.float().argmax(-1))
        assert got[0][4] and got[1][4], "TP chunked prefill differs from TP one-shot prefill"
        for rank in (0, 1):
            sl = slice(rank, rank + 1)                                   # one kv head per rank
            k0, v0 = torch.from_numpy(got[rank][2][0]), torch.from_numpy(got[rank][3][0])
            assert torch.equal(k0, kc_o[0][:, :, sl].float()) and torch.equal(v0, vc_o[0][:, :, sl].float()), \
                "layer-0 K/V of this rank's heads must be bit-identical to the oracle's"
            k1 = torch.from_numpy(got[rank][2][1])
            rel_k = ((k1 - kc_o[1][:, :, sl].float()).norm() / kc_o[1][:, :, sl].float().norm()).item()
            assert rel_k <= 2e-2, rel_k
