"""CPU tests (gloo, world_size 2) of the tensor-parallel host logic: head / weight partitioning and the row-parallel
exchange, checked against the single-rank oracle layer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import layer as OL
from oracle import ops as O
from oracle import quant as OQ
from xllm_b200 import parallel as P

BF16 = torch.bfloat16


def test_partition_heads_rules():
    # qwen2_attention.cpp:47-65
    hp = P.partition_heads(28, 4, 1, 2)
    assert (hp.num_heads, hp.num_kv_heads, hp.kv_replicas, hp.q_head0, hp.kv_head0) == (14, 2, 1, 14, 2)
    hp = P.partition_heads(28, 4, 3, 4)
    assert (hp.num_heads, hp.num_kv_heads, hp.q_head0, hp.kv_head0) == (7, 1, 21, 3)
    hp = P.partition_heads(64, 8, 5, 8)                     # Llama-3-70B TP8: 8 q / 1 kv head per GPU
    assert (hp.num_heads, hp.num_kv_heads, hp.kv_replicas, hp.kv_head0) == (8, 1, 1, 5)
    hp = P.partition_heads(32, 4, 5, 8)                     # kv heads replicated on tp / n_kv = 2 ranks
    assert (hp.num_heads, hp.num_kv_heads, hp.kv_replicas, hp.kv_head0) == (4, 1, 2, 2)
    with pytest.raises(ValueError):
        P.partition_heads(28, 4, 0, 8)                      # CHECK(total_num_heads % tp_size == 0)


def test_shard_indices_cover_everything_once():
    HQ, HKV, D, I, tp = 28, 4, 128, 18944, 4
    rows = torch.cat([P.shard_qkv_rows(HQ, HKV, D, r, tp) for r in range(tp)])
    assert sorted(rows.tolist()) == list(range((HQ + 2 * HKV) * D))
    gu = torch.cat([P.shard_gate_up_rows(I, r, tp) for r in range(tp)])
    assert sorted(gu.tolist()) == list(range(2 * I))
    cols = [P.shard_cols(I, r, tp) for r in range(tp)]
    assert cols[0].start == 0 and cols[-1].stop == I and all(cols[i].stop == cols[i + 1].start for i in range(tp - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


H, NH, NKV, D, I, GS = 256, 8, 2, 64, 512, 64


def _logical_layer(seed=2026):
    g = torch.Generator().manual_seed(seed)

    def lin(n, k, bias=False):
        w = (torch.randn(n, k, generator=g) * 0.05).to(BF16)
        q, s, z = OQ.quantize(w, 4, GS)
        return dict(q=q, s=s, z=z, w=OQ.dequantize(q, s, z, GS), b=(torch.randn(n, generator=g) * 0.05).to(BF16) if bias else None)
    return dict(qkv=lin((NH + 2 * NKV) * D, H, True), o=lin(H, NH * D), gate_up=lin(2 * I, H), down=lin(H, I),
                in_norm=(1 + 0.1 * torch.randn(H, generator=g)).to(BF16), post_norm=(1 + 0.1 * torch.randn(H, generator=g)).to(BF16))


def _inputs(seed=7):
    g = torch.Generator().manual_seed(seed)
    T = 5
    x = torch.randn(T, H, generator=g).to(BF16)
    res = torch.randn(T, H, generator=g).to(BF16)
    nblk, bs = 6, 16
    kc = torch.randn(nblk, bs, NKV, D, generator=g).to(BF16)
    vc = torch.randn(nblk, bs, NKV, D, generator=g).to(BF16)
    # one prefill request of 5 tokens written to block 2
    meta = OL.AttnMeta(True, False, torch.tensor([0, T], dtype=torch.int32), torch.tensor([0, T], dtype=torch.int32),
                       torch.arange(2 * bs, 2 * bs + T, dtype=torch.int32))
    return x, res, kc, vc, meta, torch.arange(T)


def _layer_forward(L, x, res, kc, vc, meta, pos, nh, nkv, reduce_fn):
    cs = O.compute_cos_sin_cache(D, 128, 1000000, BF16)
    attn = OL.Qwen2AttentionOracle(L["qkv"]["w"], L["qkv"]["b"], L["o"]["w"], nh, nkv, D, cs,
                                   o_linear=lambda a, w, b: reduce_fn(O.linear(a, w, b)))
    dl = OL.Qwen2DecoderLayerOracle(attn, L["in_norm"], L["post_norm"], 1e-6, lambda h: O.linear(h, L["gate_up"]["w"]),
                                    lambda h: reduce_fn(O.linear(h, L["down"]["w"])))
    return dl.forward(x, res, pos, meta, kc, vc)


def _worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pg = P.ProcessGroup()
    L = _logical_layer()
    x, res, kc, vc, meta, pos = _inputs()
    hp = P.partition_heads(NH, NKV, rank, world)
    shard = dict(
        qkv=P.shard_linear("w4", L["qkv"], P.shard_qkv_rows(NH, NKV, D, rank, world), None, GS),
        o=P.shard_linear("w4", L["o"], None, P.shard_cols(NH * D, rank, world), GS, rank),
        gate_up=P.shard_linear("w4", L["gate_up"], P.shard_gate_up_rows(I, rank, world), None, GS),
        down=P.shard_linear("w4", L["down"], None, P.shard_cols(I, rank, world), GS, rank),
        in_norm=L["in_norm"], post_norm=L["post_norm"])
    # the sharded int4 tensors dequantise to exactly the matching slice of the full dequantised weight
    for name in ("qkv", "o", "gate_up", "down"):
        assert torch.equal(OQ.dequantize(shard[name]["q"], shard[name]["s"], shard[name]["z"], GS), shard[name]["w"])

    def reduce_fn(t):                       # fp32 sum of the bf16 partials, one rounding (what csrc/allreduce.cu does)
        f = t.float()
        P.reduce(f, pg)
        return f.to(BF16)
    kv_sl = slice(hp.kv_head0, hp.kv_head0 + hp.num_kv_heads)
    y, r = _layer_forward(shard, x, res, kc[:, :, kv_sl].contiguous(), vc[:, :, kv_sl].contiguous(), meta, pos,
                          hp.num_heads, hp.num_kv_heads, reduce_fn)
    # gather (lm_head-style) sanity: ranks hold different column blocks
    g = P.gather(torch.full((2, 3), float(rank)), pg, dim=-1)
    assert g.shape == (2, 3 * world) and torch.equal(g[:, 3 * rank:3 * rank + 3], torch.full((2, 3), float(rank)))
    if rank == 0:
        out_q.put((y.float().numpy(), r.float().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_tp2_layer_matches_single_rank_oracle():
    from tests.util import assert_close_bf16
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    y_tp, r_tp = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    L = _logical_layer()
    x, res, kc, vc, meta, pos = _inputs()
    y, r = _layer_forward(L, x, res, kc, vc, meta, pos, NH, NKV, lambda t: t)
    # partial sums are rounded to bf16 per rank before the exchange: a 1-2 ulp effect on the reduced rows
    # (elements that cancel to ~0 are compared on the scale of the terms: atol = 1 ulp of an O(1) value)
    assert_close_bf16(torch.from_numpy(r_tp), r, ulps=2, rel_l2=2e-3, what="TP2 residual stream", atol=2.0 ** -7)
    assert_close_bf16(torch.from_numpy(y_tp), y, ulps=1e9, rel_l2=1e-2, what="TP2 layer output")   # MLP of a 1-ulp-perturbed input


def _subgroup_worker(rank, world, tp, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pg = P.make_tp_group(rank, world, tp)
    assert pg.world_size == tp and pg.rank == rank % tp
    # reduce / gather stay inside the replica: ranks of the other replica hold different values
    t = torch.full((3,), float(rank + 1))
    P.reduce(t, pg)
    g = P.gather(torch.full((1, 2), float(rank)), pg, dim=-1)
    out_q.put((rank, t.tolist(), g.flatten().tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_tp_x_dp_subgroups_world4():
    """bench.py --gpus 8 runs two TP4 replicas; the same group construction at world 4 = 2 replicas x TP2 on gloo."""
    world, tp = 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_subgroup_worker, args=(r, world, tp, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, red, gat in got:
        base = (rank // tp) * tp
        assert red == [float(sum(r + 1 for r in range(base, base + tp)))] * 3
        assert gat == [float(base), float(base), float(base + 1), float(base + 1)]


def test_row_parallel_bias_only_on_rank0():
    """linear.cpp:1508-1511: the bias of a row-parallel linear is added once, i.e. carried by rank 0 only, so the
    all-reduced sum of the per-rank outputs equals the unsharded linear."""
    g = torch.Generator().manual_seed(3)
    w = (torch.randn(32, 128, generator=g) * 0.05).to(BF16)
    b = (torch.randn(32, generator=g) * 0.5).to(BF16)
    x = torch.randn(4, 128, generator=g).to(BF16)
    tp = 4
    total = torch.zeros(4, 32)
    for rank in range(tp):
        sh = P.shard_linear("bf16", dict(w=w, b=b), None, P.shard_cols(128, rank, tp), 64, rank)
        assert (sh["b"] is not None) == (rank == 0)
        total += O.linear(x[:, P.shard_cols(128, rank, tp)], sh["w"], sh["b"]).float()
    ref = O.linear(x, w, b).float()
    assert (total - ref).abs().max() < 0.05 and (total - ref - b.float()).abs().max() > 0.1
    # column-parallel shards keep their own rows of the bias on every rank
    rows = torch.arange(8, 16)
    assert torch.equal(P.shard_linear("bf16", dict(w=w, b=b), rows, None, 64, 3)["b"], b[rows])


def test_fp8_linear_shards_column_and_row_parallel():
    """FP8 W8A8 logical weights (e4m3 weight, per-tensor or per-channel weight scale, static activation scale) through shard_linear:
    column-parallel shards concatenate to the full output exactly; row-parallel partials (each rounded to bf16, as every rank's
    cutlass_scaled_mm does before the all-reduce, linear.cpp:1518-1520) sum to the full output within the partials' rounding."""
    g = torch.Generator().manual_seed(9)
    N, K, tp = 96, 256, 2
    w = torch.randn(N, K, generator=g) * 0.05
    x = torch.randn(5, K, generator=g).to(BF16)
    in_scale = torch.tensor([0.02])
    for per_channel in (False, True):
        w_scale = (w.abs().amax(1) / 448.0) if per_channel else (w.abs().max() / 448.0).reshape(1)
        w8 = (w / (w_scale[:, None] if per_channel else w_scale)).clamp(-448, 448).to(torch.float8_e4m3fn)
        part = dict(w8=w8, w_scale=w_scale, in_scale=in_scale, b=None)
        full = O.fp8_linear(x, w8, w_scale, in_scale)
        cols_out = []
        for r in range(tp):
            rows = torch.arange(r * N // tp, (r + 1) * N // tp)
            sh = P.shard_linear("fp8", part, rows, None, 0, r)
            assert sh["w8"].shape == (N // tp, K) and sh["in_scale"] is in_scale
            assert sh["w_scale"].numel() == (N // tp if per_channel else 1)
            cols_out.append(O.fp8_linear(x, sh["w8"], sh["w_scale"], sh["in_scale"]))
        assert torch.equal(torch.cat(cols_out, 1), full)
        x8, _ = O.fp8_scaled_quantize(x, in_scale)
        acc = torch.zeros(5, N)
        for r in range(tp):
            sh = P.shard_linear("fp8", part, None, P.shard_cols(K, r, tp), 0, r)
            assert sh["w8"].shape == (N, K // tp)
            cs = P.shard_cols(K, r, tp)
            acc += O.fp8_scaled_matmul(x8[:, cs], sh["w8"], in_scale, sh["w_scale"]).float()
        scale = (x8.float().abs() @ w8.float().abs().t()) * in_scale * (w_scale[None, :] if per_channel else w_scale)
        assert ((acc - full.float()).abs() <= 2.0 ** -7 * scale + 1e-6).all()


def test_partition_properties_over_the_supported_head_configs():
    """for every (q heads, kv heads, tp) the reference accepts (qwen2_attention.cpp:47-65: q heads divisible by tp; kv heads either
    divisible by tp or tp divisible by kv heads, then replicated): q rows of all ranks partition the q block exactly; k / v rows
    cover every kv head, each exactly max(1, tp / kv) times; every rank's q heads map onto the kv heads it holds (GQA group intact);
    gate/up rows and row-parallel column slices partition their dimension."""
    D = 8
    for hq, hkv in ((28, 4), (64, 8), (32, 8), (16, 16), (8, 1), (12, 2)):
        for tp in (1, 2, 4, 8, 16):
            if hq % tp or not (hkv % tp == 0 or tp % hkv == 0):
                with pytest.raises(ValueError):
                    P.partition_heads(hq, hkv, 0, tp)
                continue
            group = hq // hkv
            q_rows, k_count, v_count = [], torch.zeros(hkv, dtype=torch.int64), torch.zeros(hkv, dtype=torch.int64)
            for r in range(tp):
                hp = P.partition_heads(hq, hkv, r, tp)
                assert hp.num_heads == hq // tp and hp.num_kv_heads == max(1, hkv // tp) and hp.kv_replicas == max(1, tp // hkv)
                rows = P.shard_qkv_rows(hq, hkv, D, r, tp)
                nq, nkv = hp.num_heads * D, hp.num_kv_heads * D
                q, k, v = rows[:nq], rows[nq:nq + nkv], rows[nq + nkv:]
                assert len(v) == nkv and (q < hq * D).all() and ((k >= hq * D) & (k < (hq + hkv) * D)).all() and (v >= (hq + hkv) * D).all()
                q_rows.append(q)
                kh = ((k - hq * D) // D).unique()
                vh = ((v - (hq + hkv) * D) // D).unique()
                assert torch.equal(kh, vh) and kh.tolist() == list(range(hp.kv_head0, hp.kv_head0 + hp.num_kv_heads))
                k_count[kh] += 1
                v_count[vh] += 1
                # GQA: the kv head of every local q head is one this rank holds
                qh = (q // D).unique()
                assert set((qh // group).tolist()) <= set(kh.tolist())
            assert sorted(torch.cat(q_rows).tolist()) == list(range(hq * D))
            assert (k_count == max(1, tp // hkv)).all() and (v_count == max(1, tp // hkv)).all()
    for inter, tp in ((18944, 4), (28672, 8), (96, 2)):
        gu = torch.cat([P.shard_gate_up_rows(inter, r, tp) for r in range(tp)])
        assert sorted(gu.tolist()) == list(range(2 * inter))
        for r in range(tp):                                               # a rank's gate rows and up rows are the SAME intermediate columns
            rows = P.shard_gate_up_rows(inter, r, tp)
            half = len(rows) // 2
            assert torch.equal(rows[:half] + inter, rows[half:])
            cs = P.shard_cols(inter, r, tp)
            assert rows[0].item() == cs.start and rows[half - 1].item() == cs.stop - 1
