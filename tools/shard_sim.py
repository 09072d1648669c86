"""Development aid: the per-GPU compute of the Llama-3-70B FP8 scale target at TP = N, run on ONE GPU as a model whose
shapes are the rank-0 shard (n_heads / N, n_kv_heads / N, intermediate / N, vocab / N; hidden stays) - everything a rank
does per decode step except the two exchanges per layer.  Prints the step time; run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel breakdown.
  python tools/shard_sim.py 8 [batch=32] [ctx=8192]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xllm_b200.qwen2 import Qwen2Config, Qwen2DecodeRunner, Qwen2Weights  # noqa: E402


def main():
    tp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
    layers = int(os.environ.get("XB_SIM_LAYERS", "80"))
    cfg = Qwen2Config.llama3_70b(n_heads=64 // tp, n_kv_heads=max(1, 8 // tp), intermediate_size=28672 // tp,
                                 vocab_size=128256 // tp, num_layers=layers)
    dev = "cuda"
    w = Qwen2Weights.synthetic(cfg, dev, seed=1)
    r = Qwen2DecodeRunner(cfg, w, max_batch=batch, max_ctx=ctx, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    for li in range(cfg.num_layers):
        r.k_caches[li].normal_(generator=g)
        r.v_caches[li].normal_(generator=g)
    bs = cfg.block_size
    npg = ctx // bs
    perm = (torch.randperm(r.num_blocks - 1) + 1).tolist()
    pages, indptr, slots = [], [0], []
    for b in range(batch):
        pb = perm[b * npg:(b + 1) * npg]
        pages += pb
        indptr.append(len(pages))
        slots.append(pb[(ctx - 1) // bs] * bs + (ctx - 1) % bs)
    r.set_inputs_host(list(range(100, 100 + batch)), [ctx - 1] * batch, slots, indptr, pages, [(ctx - 1) % bs + 1] * batch)
    r.step()
    r.capture()
    for _ in range(3):
        r.run_device_only()
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 5
    for _ in range(n):
        r.run_device_only()
    b_.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b_) / n
    step_bytes = w.weight_bytes() + cfg.num_layers * 2 * batch * ctx * r.nkv * cfg.head_dim * 2
    print(f"TP{tp} shard on one GPU: batch {batch} ctx {ctx} layers {layers}: {ms:.3f} ms/step  ({step_bytes / 1e9:.1f} GB -> "
          f"{step_bytes / ms / 1e6:.0f} GB/s, {step_bytes / ms / 1e6 / 6482.4:.1%} of HBM peak)  plan: chunk {r.plan.chunk_tokens} "
          f"splits {r.plan.max_splits}", flush=True)


if __name__ == "__main__":
    main()
