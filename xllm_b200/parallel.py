"""Tensor parallelism of the hot path: weight partitioning and the row-parallel exchange.

Mirrors the reference's TP rules exactly:
  * heads: n_heads / tp per rank; kv heads n_kv / tp, or ONE kv head replicated on tp / n_kv ranks when n_kv < tp
    (xllm/core/layers/common/qwen2_attention.cpp:47-65)
  * qkv_proj, gate_up_proj: column parallel (shard `out`), each fused part sharded separately
    (layers/common/linear.cpp:523-614, 1084-1146); o_proj, down_proj: row parallel (shard `in`), partial sums
    all-reduced (linear.cpp:1405-1522, reduce at :1518-1520 -> parallel_state.cpp:183-192)
  * lm_head: column parallel with gather_output (linear.cpp:712-714 -> parallel_state.cpp:89-102)
One process per GPU; torch.distributed is the plumbing (rendezvous, NCCL / gloo groups, symmetric-memory handles).
The exchange itself is either NCCL all-reduce (any size, the correctness baseline) or this library's one-shot kernels
over NVLink peer memory fused with the following RMSNorm (decode-sized messages): see csrc/allreduce.cu.
"""
import ctypes
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist

BF16 = torch.bfloat16


# ---------------------------------------------------------------------------------------------------------------
# partitioning arithmetic (pure integer / slicing logic: exercised on CPU with gloo in tests/test_parallel_cpu.py)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class HeadPartition:
    num_heads: int          # q heads on this rank
    num_kv_heads: int       # kv heads on this rank
    kv_replicas: int        # ranks sharing one kv head
    q_head0: int            # first global q head
    kv_head0: int           # first global kv head


def partition_heads(n_heads: int, n_kv_heads: int, rank: int, tp: int) -> HeadPartition:
    """qwen2_attention.cpp:47-65."""
    if n_heads % tp != 0:
        raise ValueError(f"n_heads {n_heads} not divisible by tp {tp}")
    nh = n_heads // tp
    if n_kv_heads >= tp:
        if n_kv_heads % tp != 0:
            raise ValueError(f"n_kv_heads {n_kv_heads} not divisible by tp {tp}")
        nkv, rep = n_kv_heads // tp, 1
        kv0 = rank * nkv
    else:
        if tp % n_kv_heads != 0:
            raise ValueError(f"tp {tp} not divisible by n_kv_heads {n_kv_heads}")
        nkv, rep = 1, tp // n_kv_heads
        kv0 = rank // rep
    return HeadPartition(nh, nkv, rep, rank * nh, kv0)


def shard_qkv_rows(n_heads, n_kv_heads, head_dim, rank, tp) -> torch.Tensor:
    """row indices of the fused [q | k | v] projection owned by `rank` (column-parallel, per part)."""
    hp = partition_heads(n_heads, n_kv_heads, rank, tp)
    q = torch.arange(hp.q_head0 * head_dim, (hp.q_head0 + hp.num_heads) * head_dim)
    k0 = n_heads * head_dim
    k = k0 + torch.arange(hp.kv_head0 * head_dim, (hp.kv_head0 + hp.num_kv_heads) * head_dim)
    v = k + n_kv_heads * head_dim
    return torch.cat([q, k, v])


def shard_gate_up_rows(intermediate, rank, tp) -> torch.Tensor:
    """gate and up are sharded separately and re-fused (linear.cpp fused-column load: gate_proj. / up_proj.)."""
    if intermediate % tp != 0:
        raise ValueError("intermediate_size not divisible by tp")
    per = intermediate // tp
    g = torch.arange(rank * per, (rank + 1) * per)
    return torch.cat([g, intermediate + g])


def shard_cols(in_features, rank, tp) -> slice:
    """row-parallel input slice."""
    if in_features % tp != 0:
        raise ValueError("in_features not divisible by tp")
    per = in_features // tp
    return slice(rank * per, (rank + 1) * per)


def shard_linear(kind: str, part: dict, rows: Optional[torch.Tensor], cols: Optional[slice], group_size: int,
                 rank: int = 0) -> dict:
    """Shard one logical linear (dict with w | (q, s, z) and optional b).  `rows` selects output rows (column
    parallel), `cols` selects input columns (row parallel; must be aligned to the quantisation group).  `rank` decides
    who carries the bias of a ROW-parallel shard: rank 0 only, so that it is added once after the all-reduce
    (linear.cpp:1508-1511)."""
    out = {}
    if "q" in part:
        q, s, z = part["q"], part["s"], part["z"]
        if rows is not None:
            q, s, z = q[rows], s[rows], z[rows]
        if cols is not None:
            if cols.start % group_size or cols.stop % group_size:
                raise ValueError("row-parallel shard must align with the quantisation group")
            q = q[:, cols]
            s = s[:, cols.start // group_size:cols.stop // group_size]
            z = z[:, cols.start // group_size:cols.stop // group_size]
        out.update(q=q.contiguous(), s=s.contiguous(), z=z.contiguous())
    w = part.get("w")
    if w is not None:
        if rows is not None:
            w = w[rows]
        if cols is not None:
            w = w[:, cols]
        out["w"] = w.contiguous()
    w8 = part.get("w8")
    if w8 is not None:
        # FP8 W8A8 (fp8 linear, linear.cpp:137-182): e4m3 weight [N, K] sliced like a bf16 one; a per-tensor weight scale and the
        # static activation scale are shared by every shard, a per-channel weight scale [N] follows the output rows
        if rows is not None:
            w8 = w8[rows]
        if cols is not None:
            w8 = w8[:, cols]
        ws = part["w_scale"]
        if rows is not None and ws.numel() > 1:
            ws = ws.reshape(-1)[rows]
        out.update(w8=w8.contiguous(), w_scale=ws.contiguous(), in_scale=part.get("in_scale"))
    b = part.get("b")
    if b is not None:
        # bias is added once: column-parallel shards carry their rows; row-parallel only on rank 0 (linear.cpp:1508-1511)
        if rows is not None:
            out["b"] = b[rows].contiguous()
        elif cols is not None:
            out["b"] = b if rank == 0 else None
        else:
            out["b"] = b
    else:
        out["b"] = None
    return out


# ---------------------------------------------------------------------------------------------------------------
# process group wrapper (ProcessGroup, framework/parallel_state/process_group.h:41-139)
# ---------------------------------------------------------------------------------------------------------------
class ProcessGroup:
    def __init__(self, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world_size = dist.get_world_size(self.group)

    def allreduce(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t, group=self.group)
        return t

    def allgather_base(self, t: torch.Tensor) -> torch.Tensor:
        t = t.contiguous()
        out = torch.empty((self.world_size * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out.view((self.world_size,) + tuple(t.shape))


def make_tp_group(rank: int, world: int, tp: int) -> ProcessGroup:
    """TP x DP layout of one node: `world // tp` replicas of `tp` consecutive ranks (ParallelArgs dp_size / tp groups,
    framework/parallel_state/parallel_args.h).  dp == 1 uses the default group; otherwise EVERY rank creates every
    sub-group in the same order (a c10d requirement) and keeps its own."""
    if tp < 1 or world % tp != 0:
        raise ValueError(f"world {world} is not a multiple of tp {tp}")
    dp = world // tp
    if dp == 1:
        return ProcessGroup()
    mine = None
    for gidx in range(dp):
        ranks = list(range(gidx * tp, (gidx + 1) * tp))
        grp = dist.new_group(ranks)
        if rank in ranks:
            mine = grp
    return ProcessGroup(mine)


def reduce(t: torch.Tensor, pg: Optional[ProcessGroup]) -> torch.Tensor:
    """parallel_state::reduce (parallel_state.cpp:183-192): in-place sum all-reduce."""
    if pg is None or pg.world_size == 1:
        return t
    return pg.allreduce(t)


def gather(t: torch.Tensor, pg: Optional[ProcessGroup], dim: int = -1) -> torch.Tensor:
    """parallel_state::gather (parallel_state.cpp:89-102): all-gather and concatenate along `dim`."""
    if pg is None or pg.world_size == 1:
        return t
    stacked = pg.allgather_base(t)
    return torch.cat(stacked.unbind(0), dim=dim).contiguous()


# ---------------------------------------------------------------------------------------------------------------
# NVLink peer-memory exchange (CUDA only)
# ---------------------------------------------------------------------------------------------------------------
class PeerExchange:
    """Symmetric buffers + signal pads for the fused all-reduce kernels.  Two data buffers (A: o_proj, B: down_proj)
    alternate so no trailing barrier is needed (see csrc/allreduce.cu)."""

    MAX_CTAS = 64

    def __init__(self, pg: ProcessGroup, max_tokens: int, hidden: int, device):
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import lib
        self.pg, self.hidden, self.max_tokens = pg, hidden, max_tokens
        self.lib = lib()
        n = max_tokens * hidden
        try:                                   # older torch needs the group enabled explicitly; newer ones do it lazily
            symm_mem.enable_symm_mem_for_group(pg.group.group_name)
        except Exception:
            pass
        self.bufs, self.handles = [], []
        for _ in range(2):
            t = symm_mem.empty(n, dtype=BF16, device=device)
            h = symm_mem.rendezvous(t, pg.group.group_name)
            self.bufs.append(t)
            self.handles.append(h)
        flags = symm_mem.empty(self.MAX_CTAS * 8, dtype=torch.int32, device=device)
        flags.zero_()
        self.flags = flags
        self.flags_h = symm_mem.rendezvous(flags, pg.group.group_name)
        self.epoch = torch.zeros(self.MAX_CTAS, dtype=torch.int32, device=device)
        W = pg.world_size
        self._data_ptrs = [(ctypes.c_void_p * W)(*[int(p) for p in h.buffer_ptrs]) for h in self.handles]
        self._flag_ptrs = (ctypes.c_void_p * W)(*[int(p) for p in self.flags_h.buffer_ptrs])
        torch.cuda.synchronize()
        dist.barrier(group=pg.group)

    def partial_buffer(self, which: int, tokens: int) -> torch.Tensor:
        """where the row-parallel linear of this rank writes its partial [tokens, hidden]."""
        return self.bufs[which][: tokens * self.hidden].view(tokens, self.hidden)

    def allreduce_add_rms_norm(self, which, out, residual, weight, eps, tokens):
        from ._lib import check
        rc = self.lib.xb_allreduce_add_rms_norm_bf16(
            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(residual.data_ptr()), ctypes.c_void_p(weight.data_ptr()),
            self._data_ptrs[which], self._flag_ptrs, ctypes.c_void_p(self.epoch.data_ptr()), ctypes.c_int(self.pg.rank),
            ctypes.c_int(self.pg.world_size), ctypes.c_float(eps), ctypes.c_int(tokens), ctypes.c_int(self.hidden),
            ctypes.c_int(self.MAX_CTAS), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        check(rc, "allreduce_add_rms_norm")

    def allreduce(self, which, out, tokens):
        from ._lib import check
        rc = self.lib.xb_oneshot_allreduce_bf16(
            ctypes.c_void_p(out.data_ptr()), self._data_ptrs[which], self._flag_ptrs, ctypes.c_void_p(self.epoch.data_ptr()),
            ctypes.c_int(self.pg.rank), ctypes.c_int(self.pg.world_size), ctypes.c_int64(tokens * self.hidden),
            ctypes.c_int(self.MAX_CTAS), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        check(rc, "oneshot_allreduce")
