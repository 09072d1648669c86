"""Launches ONE target kernel a few times on BASELINE configs[1] shapes so that `ncu --set full -k regex:<name> -s 2 -c 1`
captures a warmed-up launch (see profiles/ for the summaries made from these captures).
  python tools/profile_targets.py gate_up     W4A16 gate_up GEMV + SiLU*mul epilogue (37888 x 3584): the step's dominant launch
  python tools/profile_targets.py down        W4A16 down GEMV (3584 x 18944)
  python tools/profile_targets.py qkv         W4A16 qkv GEMV + RoPE / KV-scatter epilogue (4608 x 3584)
  python tools/profile_targets.py decode      paged decode attention, batch 1, ctx 4096, 28 / 4 heads of 128
  python tools/profile_targets.py decode64    paged decode attention, batch 64, ctx 4096
  python tools/profile_targets.py gemm_w4 | gemm_w4_pair | gemm_bf16_pair   tcgen05 GEMM 8192 x 4608 x 3584 (single CTA / CTA pair)
  python tools/profile_targets.py prefill | prefill_v2    causal ragged prefill attention, 4 x 2048 tokens, 28 / 4 heads of 128
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xllm_b200 import ops  # noqa: E402

DEV, BF16 = "cuda", torch.bfloat16


def w4(N, K, gs=128):
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (N // 16, K // 64, 32, 4), dtype=torch.int32, device=DEV)
    s = (torch.rand(K // gs, N, device=DEV) * 0.01 + 0.001).to(BF16).view(torch.int16).to(torch.int32) & 0xFFFF
    return qw, (s | (0x4308 << 16)).contiguous()


def main(which, reps=4):
    H = 3584
    if which in ("gate_up", "qkv"):
        # the launches of the shipped 200-launch decode step (Qwen2DecodeRunner.launch_step)
        N = 37888 if which == "gate_up" else 4608
        ws = [w4(N, H) for _ in range(reps)]
        x = torch.randn(1, H, device=DEV, dtype=BF16)
        if which == "gate_up":
            y = torch.empty(1, N // 2, device=DEV, dtype=BF16)
            fn = lambda i: ops.w4a16_gate_up_act(x, ws[i][0], ws[i][1], 128, "silu", None, y)
        else:
            y = torch.empty(1, N, device=DEV, dtype=BF16)
            pos = torch.full((1,), 4095, dtype=torch.int64, device=DEV)
            cs = torch.randn(8192, 128, device=DEV, dtype=BF16)
            slots = torch.full((1,), 300, dtype=torch.int32, device=DEV)
            kc = torch.zeros(8, 128, 4, 128, device=DEV, dtype=BF16)
            vc = torch.zeros_like(kc)
            fn = lambda i: ops.w4a16_decode_fused(x, ws[i][0], ws[i][1], 128, None, y, epilogue="rope_cache", positions=pos,
                                                  cos_sin_cache=cs, slot_ids=slots, key_cache=kc, value_cache=vc, num_heads=28,
                                                  num_kv_heads=4, head_dim=128)
    elif which == "down":
        K = 18944
        ws = [w4(H, K) for _ in range(reps)]
        x = torch.randn(1, K, device=DEV, dtype=BF16)
        y = torch.empty(1, H, device=DEV, dtype=BF16)
        fn = lambda i: ops.w4a16_linear_small_m(x, ws[i][0], ws[i][1], 128, None, y)
    elif which in ("decode", "decode64"):
        B = 1 if which == "decode" else 64
        HQ, HKV, D, page, ctx = 28, 4, 128, 128, 4096
        npg = ctx // page
        nblocks = B * npg + 1
        caches = [(torch.randn(nblocks, page, HKV, D, device=DEV, dtype=BF16), torch.randn(nblocks, page, HKV, D, device=DEV, dtype=BF16))
                  for _ in range(min(reps, 2 if B > 1 else reps))]
        q = torch.randn(B, HQ, D, device=DEV, dtype=BF16)
        out = torch.empty_like(q)
        indptr = torch.arange(0, (B + 1) * npg, npg, dtype=torch.int32, device=DEV)
        indices = (torch.randperm(nblocks - 1, device=DEV) + 1).to(torch.int32)
        last = torch.full((B,), page, dtype=torch.int32, device=DEV)
        plan = ops.DecodePlan(B, HQ, HKV, D, page, npg, DEV)
        fn = lambda i: ops.batch_decode(plan, q, caches[i % len(caches)][0], caches[i % len(caches)][1], indptr, indices, last,
                                        1 / math.sqrt(D), out)
    elif which in ("gemm_w4", "gemm_w4_pair", "gemm_bf16_pair"):
        # prefill GEMM of the qkv projection at a chunk of 8192 tokens (kept small: ncu replays every launch ~40 times)
        M, N, K = 8192, 4608, 3584
        ops.set_gemm_cta_pair(2 if which.endswith("pair") else 1)
        a = torch.randn(M, K, device=DEV, dtype=BF16)
        y = torch.empty(M, N, device=DEV, dtype=BF16)
        if which == "gemm_bf16_pair":
            wt = torch.randn(N, K, device=DEV, dtype=BF16) * 0.02
            fn = lambda i: ops.gemm_bf16(a, wt, None, y)
        else:
            qw, meta = w4(N, K)
            fn = lambda i: ops.gemm_w4a16(a, qw, meta, 128, None, y)
    elif which in ("prefill", "prefill_v2"):
        ops.set_prefill_variant(1 if which == "prefill_v2" else 0)
        HQ, HKV, D, S, nreq = 28, 4, 128, 2048, 4
        T = S * nreq
        qkv = torch.randn(T, (HQ + 2 * HKV) * D, device=DEV, dtype=BF16)
        cu = torch.arange(0, T + 1, S, dtype=torch.int32, device=DEV)
        o = torch.empty(T, HQ, D, device=DEV, dtype=BF16)
        q = qkv[:, :HQ * D].view(T, HQ, D)
        k = qkv[:, HQ * D:(HQ + HKV) * D].view(T, HKV, D)
        v = qkv[:, (HQ + HKV) * D:].view(T, HKV, D)
        fn = lambda i: ops.batch_prefill(q, k, v, cu, cu, 1 / math.sqrt(D), o, None, max_qo_len=S)
    else:
        raise SystemExit(__doc__)
    for i in range(reps):
        fn(i)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gate_up")
