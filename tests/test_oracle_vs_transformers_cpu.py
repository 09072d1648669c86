"""The oracle's WHOLE-MODEL composition (tests/model_parity.oracle_prefill: embedding -> [RMSNorm -> qkv (+bias) -> NeoX RoPE -> KV
scatter -> causal GQA attention -> o_proj -> add+RMSNorm -> SwiGLU MLP] x L -> final norm -> lm_head; the restatement of
llm_model_base.h:60-131 / qwen2_decoder_layer.cpp:64-112) against the public implementation of the model family the checkpoints are
published for: transformers' Qwen2ForCausalLM / LlamaForCausalLM (bf16, eager attention, CPU) loaded with the same random weights.
Pins RoPE convention, GQA head mapping, norm / residual placement, gate|up order and bias handling; the bf16 rounding order differs
between the two stacks, hence a relative-L2 bar and the same greedy tokens."""
import pytest
import torch

from oracle import batch as OB
from tests import model_parity as MP
from xllm_b200.qwen2 import Qwen2Config

BF16 = torch.bfloat16
tf = pytest.importorskip("transformers")


def _hf_state(cfg, W):
    qs, kvs, inter = cfg.q_size, cfg.kv_size, cfg.intermediate_size
    sd = {"model.embed_tokens.weight": W["embed"], "model.norm.weight": W["final_norm"], "lm_head.weight": W["lm_head"]}
    for i, L in enumerate(W["layers"]):
        p = f"model.layers.{i}."
        w, b = L["qkv"]["w"], L["qkv"]["b"]
        for name, sl in (("q", slice(0, qs)), ("k", slice(qs, qs + kvs)), ("v", slice(qs + kvs, qs + 2 * kvs))):
            sd[p + f"self_attn.{name}_proj.weight"] = w[sl]
            if b is not None:
                sd[p + f"self_attn.{name}_proj.bias"] = b[sl]
        sd[p + "self_attn.o_proj.weight"] = L["o"]["w"]
        sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"] = L["gate_up"]["w"][:inter], L["gate_up"]["w"][inter:]
        sd[p + "mlp.down_proj.weight"] = L["down"]["w"]
        sd[p + "input_layernorm.weight"], sd[p + "post_attention_layernorm.weight"] = L["input_norm"], L["post_norm"]
    return sd


@pytest.mark.parametrize("family", ["qwen2", "llama"])
def test_oracle_prefill_matches_transformers(family):
    qwen = family == "qwen2"
    cfg = Qwen2Config(hidden_size=64, num_layers=2, n_heads=4, n_kv_heads=2, head_dim=16, intermediate_size=96, vocab_size=128,
                      max_position_embeddings=64, block_size=4, quant="bf16", qkv_bias=qwen, rope_theta=1000000.0 if qwen else 500000.0,
                      rms_norm_eps=1e-6 if qwen else 1e-5, name=f"tiny-{family}")
    W, _, _, _ = MP.build_case(cfg, 1, [1], seed=7)
    common = dict(vocab_size=128, hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=4,
                  num_key_value_heads=2, max_position_embeddings=64, rms_norm_eps=cfg.rms_norm_eps, rope_theta=float(cfg.rope_theta),
                  tie_word_embeddings=False, attention_dropout=0.0)
    try:
        if qwen:
            hf_cfg = tf.Qwen2Config(use_sliding_window=False, **common)
            hf_cfg._attn_implementation = "eager"
            model = tf.Qwen2ForCausalLM(hf_cfg)
        else:
            hf_cfg = tf.LlamaConfig(attention_bias=False, mlp_bias=False, **common)
            hf_cfg._attn_implementation = "eager"
            model = tf.LlamaForCausalLM(hf_cfg)
        model = model.to(BF16).eval()
        res = model.load_state_dict(_hf_state(cfg, W), strict=False)
    except Exception as e:                                                 # constructor surface differs between transformers versions
        pytest.skip(f"transformers model not constructible here: {e}")
    assert not res.missing_keys and not res.unexpected_keys, res
    lens, blocks = [5, 9, 1], [[3, 1], [6, 2, 5], [4]]
    g = torch.Generator().manual_seed(3)
    toks = [torch.randint(0, cfg.vocab_size, (n,), generator=g).tolist() for n in lens]
    meta = OB.build_paged_meta([OB.SeqState(b, 0, n) for b, n in zip(blocks, lens)], cfg.block_size)
    mk = lambda: [torch.zeros(8, cfg.block_size, cfg.n_kv_heads, cfg.head_dim, dtype=BF16) for _ in range(cfg.num_layers)]
    ours = MP.oracle_prefill(cfg, W, mk(), mk(), [t for s in toks for t in s], meta, chunked=False)
    with torch.no_grad():
        theirs = torch.stack([model(input_ids=torch.tensor([s])).logits[0, -1] for s in toks])
    rel = ((ours.float() - theirs.float()).norm() / theirs.float().norm()).item()
    assert rel <= 1e-2, f"{family}: oracle vs transformers logits rel-L2 {rel:.3e}"
    assert torch.equal(ours.float().argmax(-1), theirs.float().argmax(-1))
