"""CPU test (gloo, world_size 2) of the FusedMoE layer composition (xllm_b200/moe.py, mirroring layers/cuda/fused_moe.cpp:28-117) under
expert parallelism (experts split over the ranks, foreign experts contribute zero, all-reduce over the EP group) and under MoE tensor
parallelism (the intermediate dimension of every expert split, all-reduce over the TP group), with the library ops replaced by the
oracle.  Checked against the single-rank oracle on the magnitude the per-rank bf16 roundings scale with."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import moe as OM
from oracle import ops as O
from tests.test_parallel_cpu import _free_port
from xllm_b200 import moe as M
from xllm_b200 import parallel as P

BF16 = torch.bfloat16
E, H, I, T, K = 8, 64, 96, 9, 2


class OracleMoeOps:
    @staticmethod
    def matmul(a, b, bias=None, out=None):
        return O.linear(a, b, bias)

    @staticmethod
    def moe_fused_topk(gating_output, topk, renormalize, correction_bias=None, scoring_func="softmax"):
        return OM.moe_fused_topk(gating_output, topk, renormalize, correction_bias, scoring_func)

    @staticmethod
    def cutlass_fused_moe(x, ids, scales, w13, w2, ep_size=1, ep_rank=0, out=None):
        return OM.fused_moe(x, ids, scales, w13, w2, expert_begin=ep_rank * w13.size(0))


def _weights():
    g = torch.Generator().manual_seed(21)
    gate = (torch.randn(E, H, generator=g) * 0.3).to(BF16)
    up = (torch.randn(E, I, H, generator=g) * 0.1).to(BF16)
    gt = (torch.randn(E, I, H, generator=g) * 0.1).to(BF16)
    down = (torch.randn(E, H, I, generator=g) * 0.1).to(BF16)
    x = torch.randn(T, H, generator=g).to(BF16)
    return gate, up, gt, down, x


def _worker(rank, world, port, mode, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M.ops = OracleMoeOps
    gate, up, gt, down, x = _weights()
    pg = P.ProcessGroup()
    if mode == "ep":                                                       # experts e // (E / world) == rank live here
        el = E // world
        sl = slice(rank * el, (rank + 1) * el)
        layer = M.FusedMoE(gate, torch.cat([up[sl], gt[sl]], 1), down[sl], K, ep_size=world, ep_rank=rank, ep_pg=pg)
    else:                                                                  # every expert's intermediate columns split
        il = I // world
        cs = slice(rank * il, (rank + 1) * il)
        layer = M.FusedMoE(gate, torch.cat([up[:, cs], gt[:, cs]], 1), down[:, :, cs].contiguous(), K, tp_pg=pg)
    y = layer.forward(x)
    out_q.put((rank, y.float().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _run(mode):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return got


def test_fused_moe_layer_expert_parallel_and_tensor_parallel():
    gate, up, gt, down, x = _weights()
    scales, ids = OM.moe_fused_topk(O.linear(x, gate, None), K, True, None, "softmax")
    ref, mag = OM.fused_moe(x, ids, scales, torch.cat([up, gt], 1), down, return_abs=True)
    for mode in ("ep", "tp"):
        got = _run(mode)
        y0, y1 = torch.from_numpy(got[0]), torch.from_numpy(got[1])
        assert torch.equal(y0, y1), f"{mode}: ranks disagree after the all-reduce"
        # per-rank partials are rounded to bf16 before the exchange: a few bf16 roundings on the scale of sum_k scale_k |y2_k|
        # (TP: also of the intermediate activation, which each rank rounds for its own columns)
        bound_mag = mag
        if mode == "tp":
            # the ranks' partial y2 (over half of the intermediate columns each) can cancel in the sum: their roundings scale
            # with the partials' own magnitudes, not with the magnitude of the total
            bound_mag = torch.zeros_like(mag)
            for r in range(2):
                cs = slice(r * I // 2, (r + 1) * I // 2)
                _, m_r = OM.fused_moe(x, ids, scales, torch.cat([up[:, cs], gt[:, cs]], 1), down[:, :, cs].contiguous(), return_abs=True)
                bound_mag += m_r
        assert ((y0 - ref.float()).abs() <= 2.0 ** -6 * bound_mag + 1e-6).all(), mode
        rel = ((y0 - ref.float()).norm() / ref.float().norm()).item()
        assert rel <= 1e-2, (mode, rel)
