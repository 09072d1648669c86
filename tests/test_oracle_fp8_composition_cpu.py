"""CPU check of the oracle's FP8 layer composition (oracle/layer.py): with static input scales the norms emit e4m3 directly
(apply_norm -> RMSNormImpl::forward_fp8, qwen2_decoder_layer.cpp:64-84, rms_norm.cpp:94-128) and the linear skips its own
quantisation (linear.cpp:150-157).  Against the unfused composition (norm -> bf16 -> static quant) the e4m3 activations may
differ only where the extra bf16 rounding crosses an e4m3 boundary, the residual stream not at all in the first layer."""
import torch

from oracle import ops as O
from tests import model_parity as MP
from xllm_b200.qwen2 import Qwen2Config


def _cfg():
    return Qwen2Config(hidden_size=128, num_layers=2, n_heads=4, n_kv_heads=2, head_dim=32, intermediate_size=256,
                       vocab_size=256, block_size=16, quant="fp8", qkv_bias=False, rope_theta=500000.0, rms_norm_eps=1e-5,
                       max_position_embeddings=512, name="tiny-fp8")


def test_norm_quant_kernels_agree_with_norm_then_quant():
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(7, 128, generator=g) * 2).to(torch.bfloat16)
    r = (torch.randn(7, 128, generator=g) * 2).to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(128, generator=g)).to(torch.bfloat16)
    s = torch.tensor([0.02])
    fused = O.rms_norm_static_fp8_quant(x, w, s, 1e-5)
    plain = O.static_scaled_fp8_quant(O.rms_norm(x, w, 1e-5), s)
    # same value up to one e4m3 step where the bf16 rounding of the unfused path crosses a boundary
    d = (fused.float() - plain.float()).abs()
    assert (d <= 0.13 * plain.float().abs() + 2 ** -9).all() and (d > 0).float().mean() < 0.1
    f8, res = O.fused_add_rms_norm_static_fp8_quant(x, r, w, s, 1e-5)
    h, res2 = O.fused_add_rms_norm(x, r, w, 1e-5)
    assert torch.equal(res, res2)
    assert torch.equal(f8.view(torch.uint8), O.static_scaled_fp8_quant(h, s).view(torch.uint8)), \
        "width-8 path rounds to bf16 before the conversion: identical to norm -> quant"


def test_step_composition_with_and_without_norm_quant():
    cfg = _cfg()
    W, kcs, vcs, meta = MP.build_case(cfg, 3, [40, 17, 3], seed=11)
    clone = lambda cs: [c.clone() for c in cs]
    a, _ = MP.oracle_step(cfg, W, clone(kcs), clone(vcs), meta, fp8_norm_quant=False)
    b, _ = MP.oracle_step(cfg, W, clone(kcs), clone(vcs), meta, fp8_norm_quant=True)
    rel = ((a.float() - b.float()).norm() / a.float().norm()).item()
    assert rel < 3e-2, rel
