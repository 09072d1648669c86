"""The `plan` entry points of the TVM-FFI modules xLLM dlopen()s, called on CPU through the real run-time boundary (tvm_ffi.load_module
of "$OPS/<uri>/<uri>.so") with the reference's argument lists (flashinfer_planinfo.cpp:318-335 decode, :168-247 prefill).  Planning
is host arithmetic over host tensors (indptr_host etc.), so it runs without a device; only the runs need the GPU (test_gpu_ffi.py)."""
import os

import pytest
import torch

tvm_ffi = pytest.importorskip("tvm_ffi")
BF16 = torch.bfloat16


def _load(kind, head_dim=128):
    from xllm_b200 import build_ffi
    ops_dir = build_ffi.build()
    dec, pre = build_ffi.uris()
    uri = [u for u in (dec if kind == "decode" else pre) if f"head_dim_qk_{head_dim}_" in u][0]
    return tvm_ffi.load_module(os.path.join(ops_dir, uri, uri + ".so"))


def _ws(n):
    return torch.empty(n, dtype=torch.uint8)


def _decode_plan(mod, pages_per_req, graph=False, HQ=28, HKV=4, page=128, D=128, fws=8 << 20, iws=1 << 20, window=-1, cap=0.0, dvo=None):
    indptr = torch.tensor([0] + list(torch.tensor(pages_per_req).cumsum(0)), dtype=torch.int32)
    e = torch.empty(0, dtype=BF16)
    return list(mod["plan"](_ws(fws), _ws(iws), _ws(1 << 10), indptr, len(pages_per_req), HQ, HKV, page, graph, window, cap, D,
                            D if dvo is None else dvo, e, e))


def test_decode_plan_through_the_ffi_boundary(built_lib):
    mod = _load("decode")
    plan = _decode_plan(mod, [6, 32, 1])
    assert len(plan) == 8 and plan[4:7] == [3, 28, 4]
    assert 0 < plan[2] <= 8 << 20 and 0 < (plan[3] & 0xffffffff) <= 1 << 20        # workspace needs fit what the caller gave
    assert mod["plan_is_replay_invariant"]() == 1
    # under CUDA-graph capture the plan must not depend on the context seen at capture (the reference replays `run` without
    # re-planning: flashinfer_attention.cpp:306-311): planning at 4 pages and at 64 pages gives the same launch geometry
    assert _decode_plan(mod, [4], graph=True) == _decode_plan(mod, [64], graph=True)
    # ... and it launches at least as many KV splits as the eager plan of a short context
    assert _decode_plan(mod, [4], graph=True)[1] >= _decode_plan(mod, [4])[1]


def test_decode_plan_rejections(built_lib):
    mod = _load("decode")
    with pytest.raises(RuntimeError, match="workspace too small"):
        _decode_plan(mod, [512], fws=1 << 10)            # 37 splits of one long request need 0.5 MB of partials
    with pytest.raises(ValueError, match="sliding window"):
        _decode_plan(mod, [4], window=128)
    with pytest.raises(ValueError, match="soft cap"):
        _decode_plan(mod, [4], cap=30.0)
    with pytest.raises(ValueError, match="head_dim"):
        _decode_plan(mod, [4], dvo=64)
    with pytest.raises(TypeError):
        mod["plan"](_ws(16), _ws(16))


def _prefill_plan(mod, q_lens, kv_lens, page=16, HQ=28, HKV=4, D=128, causal=True, disable_split=False, fws=64 << 20):
    cu = lambda v: torch.tensor([0] + list(torch.tensor(v).cumsum(0)), dtype=torch.int32)
    pages = [(n + page - 1) // page for n in kv_lens]
    return list(mod["plan"](_ws(fws), _ws(1 << 20), _ws(1 << 10), cu(q_lens), cu(pages), torch.tensor(kv_lens, dtype=torch.int32),
                            sum(q_lens), len(q_lens), HQ, HKV, page, False, D, D, causal, -1, -1, disable_split, 0))


def test_prefill_plan_through_the_ffi_boundary(built_lib):
    mod = _load("prefill")
    p = _prefill_plan(mod, [100, 257], [100, 257])
    assert p[:6] == [257, 2, 357, 28, 1, 0] and p[6] == 1 and len(p) == 7           # long q: no KV split
    # a short chunk over a long history: the planner spreads the KV over the SMs (flashinfer_planinfo.cpp:168-247 decides split_kv)
    p = _prefill_plan(mod, [16], [8192])
    assert p[:6] == [16, 1, 16, 28, 1, 0] and p[6] > 1
    assert _prefill_plan(mod, [16], [8192], disable_split=True)[6] == 1
    assert _prefill_plan(mod, [16], [8192], fws=1 << 10)[6] == 1                     # no room for partials: falls back to one split
    # decode served through the prefill module (GQA group >= 4: kernels/cuda/utils.cpp:349-367, batch_decode.cpp:43-60): one query
    # row per request -> the split-KV decode plan rides behind the six header words
    p = _prefill_plan(mod, [1, 1, 1], [700, 4096, 1], page=128, causal=False)
    assert p[:6] == [1, 3, 3, 28, 0, 1] and len(p) == 6 + 8 and p[6 + 4:6 + 7] == [3, 28, 4]
