"""Whole-decode-step parity: Qwen2DecodeRunner (CUDA, through the C ABI) vs the oracle composition.
Used by tests/test_gpu_model.py and by __graft_entry__.smoke()."""
import math

import torch

from oracle import layer as OL
from oracle import ops as O
from oracle import quant as OQ

BF16 = torch.bfloat16


def build_case(cfg, B, kv_lens, seed=2026, w_std=0.05):
    """logical weights (CPU, oracle quantiser) + prefilled KV caches + step metadata.  w_std: std of the linear weights
    (0.05 gives a per-layer gain > 1 on small hidden sizes - fine for a few layers, chaotic over 24; deep stacks use 0.02)."""
    g = torch.Generator().manual_seed(seed)
    H, I = cfg.hidden_size, cfg.intermediate_size

    def lin(n, k, bias=False):
        w = (torch.randn(n, k, generator=g) * w_std).to(BF16)
        b = (torch.randn(n, generator=g) * 0.05).to(BF16) if bias else None
        if cfg.quant == "bf16":
            return dict(w=w, b=b)
        if cfg.quant == "fp8":
            # per-tensor static W8A8 (compressed-tensors style): weight = e4m3(w / w_scale), activations e4m3(x / in_scale)
            w_scale = (w.float().abs().max() / 448.0).reshape(1)
            w8 = (w.float() / w_scale).clamp(-448, 448).to(torch.float8_e4m3fn)
            return dict(w8=w8, w_scale=w_scale, in_scale=torch.tensor([0.02]), b=b)
        q, s, z = OQ.quantize(w, 8 if cfg.quant == "w8a16" else 4, cfg.group_size)
        return dict(q=q, s=s, z=z, b=b, w=OQ.dequantize(q, s, z, cfg.group_size))

    W = dict(embed=(torch.randn(cfg.vocab_size, H, generator=g) * 0.5).to(BF16),
             final_norm=(1 + 0.1 * torch.randn(H, generator=g)).to(BF16), layers=[])
    W["lm_head"] = W["embed"] if cfg.tie_word_embeddings else (torch.randn(cfg.vocab_size, H, generator=g) * 0.05).to(BF16)
    for _ in range(cfg.num_layers):
        W["layers"].append(dict(input_norm=(1 + 0.1 * torch.randn(H, generator=g)).to(BF16),
                                post_norm=(1 + 0.1 * torch.randn(H, generator=g)).to(BF16),
                                qkv=lin(cfg.q_size + 2 * cfg.kv_size, H, cfg.qkv_bias), o=lin(H, cfg.q_size),
                                gate_up=lin(2 * I, H), down=lin(H, I)))
    bs = cfg.block_size
    npg = [(n + bs - 1) // bs for n in kv_lens]
    nblocks = sum(npg) + 3
    perm = (torch.randperm(nblocks - 1, generator=g) + 1)[:sum(npg)].to(torch.int32)
    indptr = [0]
    for n in npg:
        indptr.append(indptr[-1] + n)
    last = [(n - 1) % bs + 1 for n in kv_lens]
    # the new token sits at position kv_len-1; its slot follows sequence_kv_state.cpp:96-101
    slots = [int(perm[indptr[b] + (kv_lens[b] - 1) // bs]) * bs + (kv_lens[b] - 1) % bs for b in range(B)]
    kcs = [torch.randn(nblocks, bs, cfg.n_kv_heads, cfg.head_dim, generator=g).to(BF16) for _ in range(cfg.num_layers)]
    vcs = [torch.randn(nblocks, bs, cfg.n_kv_heads, cfg.head_dim, generator=g).to(BF16) for _ in range(cfg.num_layers)]
    tokens = torch.randint(0, cfg.vocab_size, (B,), generator=g).tolist()
    meta = dict(tokens=tokens, positions=[n - 1 for n in kv_lens], slots=slots, indptr=indptr, indices=perm.tolist(),
                last=last, nblocks=nblocks)
    return W, kcs, vcs, meta


def _olin(d):
    """oracle linear for one logical weight dict (bf16 / dequantised W4 / fp8 W8A8 static)."""
    if "w8" in d:
        return lambda x, _w=None, b=None: O.fp8_linear(x, d["w8"], d["w_scale"], d["in_scale"], d["b"])
    return lambda x, _w=None, b=None: O.linear(x, d["w"], d["b"])


def oracle_step(cfg, W, kcs, vcs, meta, fp8_norm_quant=False, trace=None):
    """reference composition on CPU: returns (logits bf16 [B, vocab], next tokens).  fp8_norm_quant: the norms in front of
    qkv_proj / gate_up_proj emit e4m3 with those linears' static input scales (the reference's apply_norm for FP8 checkpoints)."""
    B = len(meta["tokens"])
    cs = O.compute_cos_sin_cache(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, BF16)
    positions = torch.tensor(meta["positions"])
    am = OL.AttnMeta(False, False, torch.arange(B + 1, dtype=torch.int32), None, torch.tensor(meta["slots"], dtype=torch.int32),
                     torch.tensor(meta["indptr"], dtype=torch.int32), torch.tensor(meta["indices"], dtype=torch.int32),
                     torch.tensor(meta["last"], dtype=torch.int32))
    x = W["embed"][torch.tensor(meta["tokens"])]
    residual = None
    for li, L in enumerate(W["layers"]):
        attn = OL.Qwen2AttentionOracle(None, None, None, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cs,
                                       linear=_olin(L["qkv"]), o_linear=_olin(L["o"]))
        nq = fp8_norm_quant and "in_scale" in L["qkv"]
        dl = OL.Qwen2DecoderLayerOracle(attn, L["input_norm"], L["post_norm"], cfg.rms_norm_eps,
                                        _olin(L["gate_up"]), _olin(L["down"]),
                                        pre_fp8_scale=L["qkv"]["in_scale"] if nq else None,
                                        post_fp8_scale=L["gate_up"]["in_scale"] if nq else None)
        x, residual = dl.forward(x, residual, positions, am, kcs[li], vcs[li])
        if trace is not None:
            trace.append((x, residual))          # (MLP output, residual stream before it is added)
    x, _ = O.fused_add_rms_norm(x, residual, W["final_norm"], cfg.rms_norm_eps)
    logits = O.linear(x, W["lm_head"])
    return logits, logits.to(torch.float32).argmax(-1)


def oracle_prefill(cfg, W, kcs, vcs, tokens, meta, chunked):
    """reference composition of a prefill (chunked=False: ragged, no history) or chunked-prefill step on CPU.
    meta: oracle.batch.PagedMeta of the step; caches are updated in place.  Returns last-token logits [B, vocab]."""
    cs = O.compute_cos_sin_cache(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, BF16)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32)
    am = OL.AttnMeta(True, chunked, i32(meta.q_cu_seq_lens), i32(meta.kv_cu_seq_lens), i32(meta.new_cache_slots),
                     i32(meta.paged_kv_indptr), i32(meta.paged_kv_indices), i32(meta.paged_kv_last_page_len))
    x, residual = W["embed"][torch.tensor(tokens)], None
    positions = torch.tensor(meta.positions)
    for li, L in enumerate(W["layers"]):
        attn = OL.Qwen2AttentionOracle(None, None, None, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cs,
                                       linear=_olin(L["qkv"]), o_linear=_olin(L["o"]))
        dl = OL.Qwen2DecoderLayerOracle(attn, L["input_norm"], L["post_norm"], cfg.rms_norm_eps,
                                        _olin(L["gate_up"]), _olin(L["down"]))
        x, residual = dl.forward(x, residual, positions, am, kcs[li], vcs[li])
    x, _ = O.fused_add_rms_norm(x, residual, W["final_norm"], cfg.rms_norm_eps)
    last = torch.tensor(meta.q_cu_seq_lens[1:]) - 1
    return O.linear(x[last], W["lm_head"])


def upload(cfg, W, device="cuda"):
    from xllm_b200 import quant
    from xllm_b200.qwen2 import Linear, Qwen2Weights
    w = Qwen2Weights(cfg)
    w.embed = W["embed"].to(device)
    w.final_norm = W["final_norm"].to(device)
    w.lm_head = Linear(cfg.vocab_size, cfg.hidden_size, "bf16")
    w.lm_head.weight = w.embed if cfg.tie_word_embeddings else W["lm_head"].to(device)

    def mk(d, n, k, gate_up=False, qkv=False):
        l = Linear(n, k, cfg.quant, cfg.group_size)
        if qkv and cfg.quant == "w4a16":
            # decode layout: rows in rope-pair order so that RoPE + KV scatter fuse into the GEMV epilogue
            qw, meta, b = quant.pack_w4_qkv_rope(d["q"], d["s"], d["z"], cfg.n_heads, cfg.n_kv_heads, cfg.head_dim,
                                                 cfg.group_size, d["b"])
            l.qweight, l.meta, l.bias = qw.to(device), meta.to(device), (b.to(device) if b is not None else None)
            l.qkv_rope_packed = True
            return l
        if gate_up and cfg.quant == "w4a16":
            qw, meta, b = quant.pack_w4_gate_up(d["q"], d["s"], d["z"], cfg.group_size, d["b"])
            l.qweight, l.meta, l.bias = qw.to(device), meta.to(device), (b.to(device) if b is not None else None)
            l.gate_up_interleaved = True
            return l
        if cfg.quant == "bf16":
            l.weight = d["w"].to(device)
        elif cfg.quant == "fp8":
            l.weight = d["w8"].to(device)
            l.weight_scale = d["w_scale"].float().to(device)
            l.input_scale = d["in_scale"].float().to(device)
        elif cfg.quant == "w8a16":
            qw, meta = quant.pack_w8(d["q"], d["s"], d["z"], cfg.group_size)
            l.qweight, l.meta = qw.to(device), meta.to(device)
        else:
            qw, meta = quant.pack_w4(d["q"], d["s"], d["z"], cfg.group_size)
            l.qweight, l.meta = qw.to(device), meta.to(device)
        l.bias = d["b"].to(device) if d["b"] is not None else None
        return l
    H, I = cfg.hidden_size, cfg.intermediate_size
    for L in W["layers"]:
        w.layers.append(dict(input_norm=L["input_norm"].to(device), post_norm=L["post_norm"].to(device),
                             qkv=mk(L["qkv"], cfg.q_size + 2 * cfg.kv_size, H, qkv=True), o=mk(L["o"], H, cfg.q_size),
                             gate_up=mk(L["gate_up"], 2 * I, H, gate_up=True), down=mk(L["down"], H, I)))
    return w


def run_decode_parity(cfg, kv_lens, use_graph=True, fused=True, seed=2026, fuse_gemv=False, fp8_norm_quant=None):
    """returns (logits, oracle logits, tokens, oracle tokens, runner, caches).  fp8_norm_quant: None = the runner's default."""
    from xllm_b200.qwen2 import Qwen2DecodeRunner
    B = len(kv_lens)
    W, kcs, vcs, meta = build_case(cfg, B, kv_lens, seed)
    runner = Qwen2DecodeRunner(cfg, upload(cfg, W), B, max(kv_lens), num_blocks=meta["nblocks"], fused_rope_cache=fused,
                               fuse_gemv=fuse_gemv)
    if fp8_norm_quant is not None:
        runner.fp8_norm_quant = bool(fp8_norm_quant)
    for li in range(cfg.num_layers):
        runner.k_caches[li].copy_(kcs[li])
        runner.v_caches[li].copy_(vcs[li])
    runner.set_inputs_host(meta["tokens"], meta["positions"], meta["slots"], meta["indptr"], meta["indices"], meta["last"])
    if use_graph:
        runner.step()            # eager once (module load), then capture
        for li in range(cfg.num_layers):
            runner.k_caches[li].copy_(kcs[li])
            runner.v_caches[li].copy_(vcs[li])
        runner.capture()
    nxt = runner.step().clone()
    logits = runner.logits.cpu()
    ref_logits, ref_next = oracle_step(cfg, W, kcs, vcs, meta, fp8_norm_quant=runner.fp8_norm_quant)
    # KV caches after the step: the new token's rotated K and V must have been scattered bit-exactly
    # (same qkv GEMM inputs -> within tolerance; compare the untouched part exactly)
    return logits, ref_logits, nxt, ref_next, runner, (kcs, vcs)


def oracle_layer_state(cfg, W, li, h, residual, kcs, vcs, meta):
    """One decoder layer of the oracle in the runner's state convention: (h = normalised layer input, residual =
    residual stream) -> (normalised input of the next layer / final norm, updated residual stream).
    Composition per qwen2_decoder_layer.cpp:89-112 with the next layer's input add+norm folded in."""
    B = len(meta["tokens"])
    cs = O.compute_cos_sin_cache(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta, BF16)
    am = OL.AttnMeta(False, False, torch.arange(B + 1, dtype=torch.int32), None, torch.tensor(meta["slots"], dtype=torch.int32),
                     torch.tensor(meta["indptr"], dtype=torch.int32), torch.tensor(meta["indices"], dtype=torch.int32),
                     torch.tensor(meta["last"], dtype=torch.int32))
    L = W["layers"][li]
    attn = OL.Qwen2AttentionOracle(None, None, None, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim, cs,
                                   linear=_olin(L["qkv"]), o_linear=_olin(L["o"]))
    a = attn.forward(torch.tensor(meta["positions"]), h, am, kcs[li].clone(), vcs[li].clone())
    h_mid, res_mid = O.fused_add_rms_norm(a, residual, L["post_norm"], cfg.rms_norm_eps)
    down = _olin(L["down"])(O.act_and_mul(_olin(L["gate_up"])(h_mid), "silu"))
    next_w = W["layers"][li + 1]["input_norm"] if li + 1 < cfg.num_layers else W["final_norm"]
    return O.fused_add_rms_norm(down, res_mid, next_w, cfg.rms_norm_eps)


def run_layerwise_parity(cfg, kv_lens, seed=2026):
    """Eager GPU step with a per-layer trace; every layer of the oracle is then fed the GPU's own state before that
    layer, so each comparison isolates ONE decoder layer (no compounding of bf16 rounding flips across layers).
    Returns a list of (gpu_h, ref_h, gpu_res, ref_res) per layer."""
    from xllm_b200.qwen2 import Qwen2DecodeRunner
    B = len(kv_lens)
    W, kcs, vcs, meta = build_case(cfg, B, kv_lens, seed)
    runner = Qwen2DecodeRunner(cfg, upload(cfg, W), B, max(kv_lens), num_blocks=meta["nblocks"])
    for li in range(cfg.num_layers):
        runner.k_caches[li].copy_(kcs[li])
        runner.v_caches[li].copy_(vcs[li])
    runner.set_inputs_host(meta["tokens"], meta["positions"], meta["slots"], meta["indptr"], meta["indices"], meta["last"])
    for d, h in ((runner.token_ids, runner.h_token_ids), (runner.positions, runner.h_positions), (runner.slots, runner.h_slots),
                 (runner.kv_indptr, runner.h_kv_indptr), (runner.kv_indices, runner.h_kv_indices), (runner.kv_last, runner.h_kv_last)):
        d.copy_(h)
    trace = []
    runner.launch_step(trace)
    torch.cuda.synchronize()
    out = []
    res_in = W["embed"][torch.tensor(meta["tokens"])]
    h_in = O.rms_norm(res_in, W["layers"][0]["input_norm"], cfg.rms_norm_eps)
    for li in range(cfg.num_layers):
        ref_h, ref_res = oracle_layer_state(cfg, W, li, h_in, res_in, kcs, vcs, meta)
        gh, gres = trace[li][0].cpu(), trace[li][1].cpu()
        out.append((gh, ref_h, gres, ref_res))
        h_in, res_in = gh, gres          # teacher forcing: next layer sees what the GPU produced
    return out
