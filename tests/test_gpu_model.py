"""GPU parity: a whole decode step of the Qwen2 stack (CUDA graph of C-ABI launches) vs the oracle composition."""
import pytest
import torch

from tests.model_parity import run_decode_parity, run_layerwise_parity
from tests.util import assert_close_bf16

pytestmark = pytest.mark.gpu


def _small(quant):
    from xllm_b200.qwen2 import Qwen2Config
    return Qwen2Config(hidden_size=256, num_layers=3, n_heads=8, n_kv_heads=2, head_dim=64, intermediate_size=512,
                       vocab_size=1024, block_size=16, quant=quant, group_size=64, max_position_embeddings=2048,
                       name="tiny")


@pytest.mark.parametrize("quant", ["w4a16", "w8a16", "bf16"])
@pytest.mark.parametrize("use_graph,fused", [(True, True), (False, False)])
def test_decode_step_matches_oracle(quant, use_graph, fused, built_lib):
    cfg = _small(quant)
    logits, ref_logits, nxt, ref_next, runner, (kcs, vcs) = run_decode_parity(cfg, [37, 300, 1], use_graph, fused)
    # Whole step: ~30 bf16-rounded ops in sequence; a 1-ulp flip (fp32 summation order) in one op perturbs
    # everything downstream, so two correct pipelines drift apart by about one bf16 ulp per element (measured:
    # rel-L2 ~7e-3 on this 3-layer, H=256 stack).  The tight per-op bar is enforced by test_layerwise_* below and
    # the per-kernel tests; here: same greedy tokens, logits within 2e-2 relative L2.
    assert_close_bf16(logits, ref_logits, ulps=1e9, rel_l2=2e-2, what="decode-step logits")
    assert torch.equal(nxt.long().cpu()[:3], ref_next), "greedy tokens differ"
    # KV caches: the oracle step scattered the new token's (rotated) K and V into kcs/vcs in place.  Rows of old tokens
    # must be bit-identical; the 3 new rows per layer carry the upstream activation drift, so only loosely close.
    bs = cfg.block_size
    new_rows = torch.zeros(kcs[0].shape[0] * bs, dtype=torch.bool)
    from tests.model_parity import build_case
    _, _, _, meta = build_case(cfg, 3, [37, 300, 1])
    new_rows[torch.tensor(meta["slots"])] = True
    for li in range(cfg.num_layers):
        for name, got, ref in (("k", runner.k_caches[li], kcs[li]), ("v", runner.v_caches[li], vcs[li])):
            g = got.cpu().view(-1, cfg.n_kv_heads * cfg.head_dim)
            r = ref.view(-1, cfg.n_kv_heads * cfg.head_dim)
            assert torch.equal(g[~new_rows], r[~new_rows]), f"{name}_cache[{li}]: an old row changed"
            assert_close_bf16(g[new_rows], r[new_rows], ulps=1e9, rel_l2=2e-2, what=f"{name}_cache[{li}] new rows")


def test_decode_step_split_rmsnorm_variant(built_lib):
    """the 5-launches-per-layer step (add+RMSNorm split between the o / down epilogues and the qkv / gate_up prologues;
    opt-in: Qwen2DecodeRunner(fuse_gemv=True)) computes the same step."""
    cfg = _small("w4a16")
    logits, ref_logits, nxt, ref_next, runner, _ = run_decode_parity(cfg, [37, 300, 1], True, True, fuse_gemv=True)
    assert runner.fuse_gemv
    assert_close_bf16(logits, ref_logits, ulps=1e9, rel_l2=2e-2, what="decode-step logits (split RMSNorm)")
    assert torch.equal(nxt.long().cpu()[:3], ref_next), "greedy tokens differ"


def test_decode_step_mlp_norm_variant(built_lib, monkeypatch):
    """XB_FUSE_MLP_NORM=1: the post-attention add+RMSNorm rides in the gate_up GEMV's prologue; same step."""
    monkeypatch.setenv("XB_FUSE_MLP_NORM", "1")
    cfg = _small("w4a16")
    logits, ref_logits, nxt, ref_next, runner, _ = run_decode_parity(cfg, [37, 300, 1], True, True)
    assert runner.fuse_mlp_norm
    assert_close_bf16(logits, ref_logits, ulps=1e9, rel_l2=2e-2, what="decode-step logits (gate_up norm prologue)")
    assert torch.equal(nxt.long().cpu()[:3], ref_next), "greedy tokens differ"


def test_decode_step_qwen2_0_5b_shape(built_lib):
    """BASELINE configs[0] architecture (Qwen2-0.5B bf16, batch 1, ctx 128) with 2 layers to keep the oracle fast."""
    from xllm_b200.qwen2 import Qwen2Config
    cfg = Qwen2Config.qwen2_0_5b(num_layers=2, vocab_size=8192, block_size=128, max_position_embeddings=4096)
    logits, ref_logits, nxt, ref_next, _, _ = run_decode_parity(cfg, [128], True, True)
    assert_close_bf16(logits, ref_logits, ulps=1e9, rel_l2=2e-2, what="qwen2-0.5b-shape logits")
    assert torch.equal(nxt.long().cpu()[:1], ref_next)


@pytest.mark.parametrize("quant", ["w4a16", "bf16"])
def test_layerwise_teacher_forced(quant, built_lib):
    """Each decoder layer against the oracle ON THE SAME INPUT (the GPU's own previous-layer output): isolates one
    layer = ~10 bf16-rounded ops incl. attention with bf16 P (measured ~1.2e-3 on the residual stream), so the bar is
    rel-L2 <= 3e-3 on the residual stream (within 2 ulps elementwise) and <= 1e-2 on the normalised input of the next
    layer (it carries the MLP's response to 1-ulp flips of its input; measured 4e-3); the 1e-3 / 1-ulp bars are enforced
    per op in test_gpu_{elementwise,linear,decode,prefill,gemm}.py."""
    cfg = _small(quant)
    for li, (gx, rx, gres, rres) in enumerate(run_layerwise_parity(cfg, [37, 300, 1])):
        assert_close_bf16(gres, rres, ulps=2, rel_l2=3e-3, what=f"layer {li} residual stream", atol=2.0 ** -6)
        assert_close_bf16(gx, rx, ulps=1e9, rel_l2=1e-2, what=f"layer {li} normalised output")


@pytest.mark.parametrize("norm_quant", [False, True])
def test_decode_step_llama_fp8_shape(norm_quant, built_lib):
    """BASELINE configs[3] flavour at toy size: Llama-style layer (no qkv bias, GQA 8) with FP8 W8A8 per-tensor static
    linears (fp8_linear_forward, linear.cpp:137-182 -> cutlass_scaled_mm), decode batch of 3.  norm_quant: the norms in front
    of qkv_proj / gate_up_proj emit e4m3 directly (the reference's apply_norm for checkpoints with static input scales,
    qwen2_decoder_layer.cpp:64-84), oracle composed the same way."""
    from xllm_b200.qwen2 import Qwen2Config
    cfg = Qwen2Config(hidden_size=512, num_layers=2, n_heads=16, n_kv_heads=2, head_dim=64, intermediate_size=1024,
                      vocab_size=2048, block_size=16, quant="fp8", qkv_bias=False, rope_theta=500000.0, rms_norm_eps=1e-5,
                      max_position_embeddings=2048, name="tiny-llama-fp8")
    logits, ref_logits, nxt, ref_next, _, caches = run_decode_parity(cfg, [200, 33, 5], True, True, fp8_norm_quant=norm_quant)
    # e4m3 activations: a 1-ulp bf16 flip upstream can move an activation across an fp8 rounding boundary (2^-4
    # relative), so the logits bar is looser than for bf16 pipelines.  The bar is the model's OWN sensitivity: the oracle's
    # response to a 1-ulp (2^-8 relative) change of every 7th element of the three input embedding rows (measured 8e-2 on
    # this toy stack: weights of std 0.05 give every layer a gain > 1) - a GPU result inside that band is as close to the
    # oracle as the oracle is to itself under the smallest representable input change.
    from tests.model_parity import build_case, oracle_step
    W, kcs, vcs, meta = build_case(cfg, 3, [200, 33, 5], 2026)
    W2 = dict(W)
    e = W["embed"].clone()
    rows = torch.tensor(meta["tokens"])
    e[rows, ::7] = (e[rows, ::7].float() * (1 + 2.0 ** -8)).to(torch.bfloat16)
    W2["embed"] = e
    pert, _ = oracle_step(cfg, W2, [c.clone() for c in kcs], [c.clone() for c in vcs], meta, fp8_norm_quant=norm_quant)
    sens = ((pert.float() - ref_logits.float()).norm() / ref_logits.float().norm()).item()
    assert_close_bf16(logits, ref_logits, ulps=1e9, rel_l2=max(5e-2, 1.25 * sens), what=f"llama-fp8 logits (1-ulp sensitivity {sens:.2e})")
    # random-init logits are nearly flat, so the argmax may flip between near-ties: the token the GPU picked must be
    # (near-)maximal under the oracle as well
    rl = ref_logits.float()
    for b in range(3):
        top = rl[b].max().item()
        assert rl[b, int(nxt[b])].item() >= top - 0.05 * abs(top), f"request {b}: GPU token is not a near-argmax of the oracle"
