import os, sys, torch
sys.path.insert(0, "/root/repo")
from xllm_b200 import ops
DEV, BF16 = "cuda", torch.bfloat16
def t_us(fn, it=6):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / it
M = 8192
for name, (N, K) in {"qkv": (4608, 3584), "gate_up": (37888, 3584), "down": (3584, 18944)}.items():
    a = torch.randn(M, K, device=DEV, dtype=BF16)
    y = torch.empty(M, N, device=DEV, dtype=BF16)
    qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (N // 16, K // 64, 32, 4), dtype=torch.int32, device=DEV)
    sc = (torch.rand(K // 128, N, device=DEV) * 0.01 + 0.001).to(BF16).view(torch.int16).to(torch.int32) & 0xFFFF
    meta = (sc | (0x4308 << 16)).contiguous()
    fl = 2.0 * M * N * K
    print(name, "skip_convert=" + os.environ.get("XB_GEMM_DEBUG_SKIP_CONVERT", "0"), "pair_bn=" + os.environ.get("XB_GEMM_PAIR_BN", "auto"),
          f"w4 pair {fl / t_us(lambda: ops.gemm_w4a16(a, qw, meta, 128, None, y)) / 1e6:7.1f} TF/s", flush=True)
