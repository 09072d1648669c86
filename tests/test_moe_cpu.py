"""CPU: the MoE router restatement (oracle/moe.py) against the ordering the reference's own test pins
(tests/core/kernels/cuda/moe/moe_topk_test.cu:31-55 cpuTopK: value descending, ties to the smaller index) and the documented
semantics of moe_fused_topk (softmax / sigmoid + correction bias / renormalize)."""
import numpy as np
import pytest
import torch

from oracle import moe as OM


def _cpu_topk(values, k):
    """moe_topk_test.cu:31-55"""
    order = sorted(range(len(values)), key=lambda i: (-values[i], i))
    return order[:k]


def test_topk_ordering_matches_reference_cpu_topk():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(7, 64, generator=g)
    x[2, 5] = x[2, 40] = x[2].max() + 1.0            # an exact tie at the top: lower index first
    w, ids = OM.moe_fused_topk(x, 6, False, None, "softmax")
    p = torch.softmax(x.float(), dim=1)
    for t in range(7):
        assert ids[t].tolist() == _cpu_topk(p[t].tolist(), 6)
        assert np.allclose(w[t].numpy(), p[t][ids[t].long()].numpy(), rtol=1e-6)
    assert ids[2, 0] == 5 and ids[2, 1] == 40


def test_sigmoid_bias_and_renormalize_semantics():
    x = torch.tensor([[0.0, 2.0, -1.0, 2.0]])
    bias = torch.tensor([10.0, 0.0, 0.0, 0.0])
    w, ids = OM.moe_fused_topk(x, 2, True, bias, "sigmoid")
    # selection uses sigmoid + bias (expert 0 wins through its bias), the weight is the biased value minus the bias
    assert ids[0].tolist() == [0, 1]
    s = torch.sigmoid(x[0])
    raw = np.array([np.float32(np.float32(s[0] + 10.0) - 10.0), np.float32(s[1])], np.float32)
    assert np.allclose(w[0].numpy(), raw / raw.sum(), rtol=1e-6)
    assert abs(float(w.sum()) - 1.0) < 1e-6
    w2, _ = OM.moe_fused_topk(x, 2, False, None, "sigmoid")
    assert np.allclose(w2[0].numpy(), [s[1], s[3]], rtol=1e-6)         # tie between experts 1 and 3 -> 1 first


def test_fused_moe_zero_for_foreign_experts():
    g = torch.Generator().manual_seed(3)
    H, I, E = 64, 32, 4
    x = torch.randn(2, H, generator=g).to(torch.bfloat16)
    fc1 = (torch.randn(E, 2 * I, H, generator=g) * 0.1).to(torch.bfloat16)
    fc2 = (torch.randn(E, H, I, generator=g) * 0.1).to(torch.bfloat16)
    ids = torch.tensor([[0, 3], [2, 1]], dtype=torch.int32)
    sc = torch.tensor([[0.6, 0.4], [0.5, 0.5]])
    full = OM.fused_moe(x, ids, sc, fc1, fc2)
    lo = OM.fused_moe(x, ids, sc, fc1[:2], fc2[:2], expert_begin=0)
    hi = OM.fused_moe(x, ids, sc, fc1[2:], fc2[2:], expert_begin=2)
    # expert parallelism: the two halves add up (before the final rounding) to the full result
    assert torch.allclose(lo.float() + hi.float(), full.float(), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("case,tokens,hidden,experts,scoring,with_bias,route_scale,rw,eid", [
    ("sigmoid", 512, 7168, 16, "sigmoid", True, 2.5, (1.25, 1.25, 1280.0), (6.0, 7.0, 6656.0)),
    ("softmax", 512, 7168, 16, "softmax", False, 2.5, (3.16604e-14, 2.5, 1280.0), (0.0, 14.0, 4413.0)),
    ("sigmoid_topk1", 128, 1024, 8, "sigmoid", False, 1.0, (0.5, 0.5, 128.0), (0.0, 1.0, 128.0))])
def test_router_reproduces_the_reference_moe_gate_known_answers(case, tokens, hidden, experts, scoring, with_bias, route_scale, rw, eid):
    """tests/core/layers/mlu/moe_gate_test.cpp:143-268 (MoEGateTest.Sigmoid / Softmax / SigmoidTopkGroup1): gate linear over
    `seeded_tensor` inputs -> scoring -> top-2 -> renormalise -> route scale, expected min / max / sum of the routing weights and of
    the expert ids recorded from the reference's hardware (expect_tensor_stats: rtol 1e-2, atol 1e-5, tests_utils.cpp:90-117).  All
    three configurations keep every expert group (topk_group == n_group), i.e. exactly the routing xllm::kernel::cuda::moe_fused_topk
    restated in oracle/moe.py implements; the gate linear is the reference's bf16 F::linear (oracle.ops.linear)."""
    from oracle import ops as O
    from oracle import seeded
    W = seeded.seeded_tensor("moe_gate_tests.gate_proj.weight", (experts, hidden), torch.bfloat16)
    x = seeded.seeded_tensor(f"moe_gate_tests.{case}.hidden_states", (tokens, hidden), torch.bfloat16)
    bias = seeded.seeded_tensor("moe_gate_tests.e_score_correction_bias", (experts,), torch.bfloat16).float() if with_bias else None
    w, ids = OM.moe_fused_topk(O.linear(x, W, None), 2, True, bias, scoring)
    w = w * route_scale

    def close(actual, expected):
        return abs(actual - expected) <= 1e-5 + 1e-2 * abs(expected)
    for got, exp in (((w.min().item(), w.max().item(), w.double().sum().item()), rw),
                     ((float(ids.min()), float(ids.max()), float(ids.sum())), eid)):
        assert all(close(a, e) for a, e in zip(got, exp)), (case, got, exp)
    if case == "softmax":
        assert int(ids.sum()) == 4413 and abs(w.min().item() / 3.16604e-14 - 1) < 1e-4      # far inside the reference's tolerance


def test_experts_and_router_match_transformers_mixtral():
    """oracle.moe.fused_moe / moe_fused_topk against the public implementation of a gated-MoE block: transformers' MixtralExperts
    (gate_up_proj [E, 2I, H] in [gate | up] order, SiLU(gate) * up, down_proj, weighted index_add) and its softmax -> top-k ->
    renormalise routing.  The reference stacks fc1 as [up | gate] (layers/cuda/fused_moe.cpp:124-126), so the halves are swapped when
    handing the same weights to the two implementations."""
    tf = pytest.importorskip("transformers")
    try:
        from transformers.models.mixtral import modeling_mixtral as mm
        cfg = tf.MixtralConfig(hidden_size=64, intermediate_size=96, num_local_experts=8, num_experts_per_tok=2, num_hidden_layers=1,
                               num_attention_heads=4, num_key_value_heads=2, vocab_size=128)
        experts = mm.MixtralExperts(cfg).to(torch.bfloat16)
        assert tuple(experts.gate_up_proj.shape) == (8, 192, 64) and tuple(experts.down_proj.shape) == (8, 64, 96)
    except Exception as e:
        pytest.skip(f"transformers MixtralExperts not constructible here: {e}")
    g = torch.Generator().manual_seed(5)
    E, H, I, T, k = 8, 64, 96, 11, 2
    gate = (torch.randn(E, I, H, generator=g) * 0.1).to(torch.bfloat16)
    up = (torch.randn(E, I, H, generator=g) * 0.1).to(torch.bfloat16)
    down = (torch.randn(E, H, I, generator=g) * 0.1).to(torch.bfloat16)
    x = torch.randn(T, H, generator=g).to(torch.bfloat16)
    logits = torch.randn(T, E, generator=g)
    w, ids = OM.moe_fused_topk(logits, k, True, None, "softmax")
    # routing: softmax over all experts, top-k, renormalise (MixtralSparseMoeBlock / MixtralTopKRouter)
    p = torch.softmax(logits.float(), -1)
    tw, ti = torch.topk(p, k, dim=-1)
    tw = tw / tw.sum(-1, keepdim=True)
    assert torch.equal(ids.long(), ti) and torch.allclose(w, tw, rtol=1e-6, atol=1e-7)
    with torch.no_grad():
        experts.gate_up_proj.copy_(torch.cat([gate, up], 1))
        experts.down_proj.copy_(down)
        theirs = experts(x, ti, tw.to(torch.bfloat16))
    ours = OM.fused_moe(x, ids, w, torch.cat([up, gate], 1), down)
    rel = ((ours.float() - theirs.float()).norm() / theirs.float().norm()).item()
    assert rel <= 1e-2, rel
