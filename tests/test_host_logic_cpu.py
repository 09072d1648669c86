"""Host-side logic of the C ABI that needs no GPU: the decode planner's invariants over a sweep of batch / head /
page geometries, the plan flag word, and the argument validation every launcher performs before its first CUDA call
(errors come back as a non-zero code + xb_last_error(), never as a crash or a silent CPU path)."""
import ctypes
import itertools

import pytest

from xllm_b200 import _lib

c_void = ctypes.c_void_p


def _plan(lib, batch, hq, hkv, d, page, max_pages, sms=148):
    plan = (ctypes.c_int64 * 8)()
    rc = lib.xb_decode_plan(plan, batch, hq, hkv, d, page, max_pages, sms)
    return rc, plan


@pytest.mark.parametrize("batch,hq,hkv,d", [(1, 28, 4, 128), (8, 28, 4, 128), (64, 28, 4, 128), (256, 28, 4, 128),
                                            (1, 64, 8, 128), (3, 14, 2, 64), (1, 32, 32, 128), (5, 40, 8, 128)])
@pytest.mark.parametrize("page,max_pages", [(16, 256), (128, 32), (1, 4096), (48, 100), (128, 1024)])
def test_decode_plan_invariants(batch, hq, hkv, d, page, max_pages, built_lib, monkeypatch):
    monkeypatch.delenv("XB_DECODE_CHUNK", raising=False)
    monkeypatch.delenv("XB_DECODE_WARPS", raising=False)
    monkeypatch.delenv("XB_DECODE_CLUSTER", raising=False)
    lib = _lib.lib()
    rc, p = _plan(lib, batch, hq, hkv, d, page, max_pages)
    assert rc == 0, lib.xb_last_error()
    chunk, splits = p[0], p[1]
    cluster = (p[7] >> 44) & 0x1F
    max_kv = page * max_pages
    assert chunk % 16 == 0 and chunk >= 64, "nominal chunks are whole 16-token blocks of at least 64 tokens"
    assert splits >= 1 and chunk * splits >= max_kv, "the chunks cover the longest admissible request"
    assert splits == 1 or 64 * (splits - 1) < max_kv, "never more splits than 64-token chunks in the longest request"
    assert splits <= 2 * d, "split merge scratch is sized for 2*head_dim splits"
    assert 1 <= cluster <= 16 and splits % cluster == 0, "splits = clusters x cluster size, at most 16 CTAs per cluster"
    # workspaces: partial O + LSE per (request, q head, split) - sized for the cluster-less fallback; one ticket word
    # per (request, kv head, head tile)
    if splits > 1:
        assert p[2] >= batch * hq * splits * (d + 1) * 4
    group = hq // hkv
    head_tiles = (group + 15) // 16
    assert (p[3] & 0xFFFFFFFF) >= batch * hkv * head_tiles * 8
    assert (p[4], p[5], p[6]) == (batch, hq, hkv)
    assert p[7] & 0xFFFF == d and (p[7] >> 16) & 0xFFFFFF == page and (p[7] >> 40) & 0xF in (4, 8)
    # one wave: with few (request, kv head) units the planner splits the KV range to fill the SMs, never beyond them
    units = batch * hkv * head_tiles
    if units >= 148:
        assert splits == 1
    else:
        assert units * splits <= 148, f"{units} units x {splits} splits overshoots one wave"


def test_decode_plan_flags_and_env(built_lib, monkeypatch):
    lib = _lib.lib()
    rc, p = _plan(lib, 1, 28, 4, 128, 128, 32)
    assert rc == 0 and (p[3] >> 32) & 1 == 0
    assert lib.xb_decode_plan_set_flags(p, 1) == 0 and (p[3] >> 32) & 1 == 1
    low = p[3] & 0xFFFFFFFF
    assert lib.xb_decode_plan_set_flags(p, 0) == 0 and (p[3] >> 32) & 1 == 0 and p[3] & 0xFFFFFFFF == low
    assert lib.xb_decode_plan_set_flags(None, 1) != 0
    monkeypatch.setenv("XB_DECODE_CHUNK", "200")          # rounded down to whole blocks -> 192 -> 22 splits of <= 192
    rc, p = _plan(lib, 1, 28, 4, 128, 128, 32)
    assert rc == 0 and p[1] == 22 and p[0] == 192 and (p[7] >> 44) & 0x1F == 1
    monkeypatch.setenv("XB_DECODE_CLUSTER", "16")         # opt-in clusters: 22 splits = 2 clusters of 11 CTAs
    rc, p = _plan(lib, 1, 28, 4, 128, 128, 32)
    assert rc == 0 and p[1] == 2 * 11 and (p[7] >> 44) & 0x1F == 11
    monkeypatch.delenv("XB_DECODE_CLUSTER")
    monkeypatch.delenv("XB_DECODE_CHUNK")
    rc, p = _plan(lib, 1, 28, 4, 128, 128, 32)            # BASELINE configs[1]: 37 splits of 112 tokens, no clusters
    assert rc == 0 and p[1] == 37 and p[0] == 112 and (p[7] >> 44) & 0x1F == 1
    monkeypatch.setenv("XB_DECODE_WARPS", "4")
    rc, p = _plan(lib, 1, 28, 4, 128, 128, 32)
    assert rc == 0 and (p[7] >> 40) & 0xF == 4


@pytest.mark.parametrize("args", [(0, 28, 4, 128, 128, 32), (1, 28, 3, 128, 128, 32), (1, 28, 4, 256, 128, 32),
                                  (1, 28, 4, 128, 0, 32), (1, 28, 4, 128, 128, 0), (1, 28, 0, 128, 128, 32)])
def test_decode_plan_rejects_bad_geometry(args, built_lib):
    lib = _lib.lib()
    rc, _ = _plan(lib, *args)
    assert rc != 0 and len(lib.xb_last_error()) > 0


def test_launchers_validate_before_touching_the_gpu(built_lib):
    """Every call below must fail in the argument checks (no device pointer is ever dereferenced: they are NULL)."""
    lib = _lib.lib()
    i32, i64, f32 = ctypes.c_int, ctypes.c_int64, ctypes.c_float
    null = c_void(0)
    cases = {
        "small-M W4: M out of range": lambda: lib.xb_linear_w4a16_small_m(null, i64(0), null, i64(0), null, null, null, i32(65),
                                                                         i32(4608), i32(3584), i32(128), null),
        "small-M W4: N not a multiple of 16": lambda: lib.xb_linear_w4a16_small_m(null, i64(0), null, i64(0), null, null, null,
                                                                                 i32(1), i32(4600), i32(3584), i32(128), null),
        "small-M W4: group does not divide K": lambda: lib.xb_linear_w4a16_small_m(null, i64(0), null, i64(0), null, null, null,
                                                                                  i32(1), i32(4608), i32(3584), i32(96), null),
        "gate_up act: unknown activation": lambda: lib.xb_linear_w4a16_gate_up_act_small_m(null, i64(0), null, i64(0), null, null,
                                                                                            null, i32(1), i32(4608), i32(3584),
                                                                                            i32(128), i32(7), null),
        "small-M bf16: M out of range": lambda: lib.xb_linear_bf16_small_m(null, i64(0), null, i64(0), null, null, i32(99), i32(64),
                                                                           i32(64), null),
        "small-M bf16: K not a multiple of 32": lambda: lib.xb_linear_bf16_small_m(null, i64(0), null, i64(0), null, null, i32(1),
                                                                                   i32(64), i32(48), null),
        "paged decode: null plan": lambda: lib.xb_paged_decode_bf16(None, null, i64(0), i64(0), null, null, i64(0), i64(0), i64(0),
                                                                    null, null, null, null, i64(0), i64(0), null, f32(1.0), null,
                                                                    null, null),
    }
    for what, call in cases.items():
        rc = call()
        assert rc != 0, f"{what}: accepted"
        assert len(lib.xb_last_error()) > 0, f"{what}: no message"
    # M == 0 is a no-op, not an error (empty decode batch)
    assert lib.xb_linear_w4a16_small_m(null, i64(0), null, i64(0), null, null, null, i32(0), i32(4608), i32(3584), i32(128),
                                       null) == 0
    assert lib.xb_linear_bf16_small_m(null, i64(0), null, i64(0), null, null, i32(0), i32(64), i32(64), null) == 0


def test_prefill_split_kv_decision(built_lib):
    """xb_prefill_plan_splits: host arithmetic of the split-KV decision for short-q / long-kv chunked prefill."""
    lib = _lib.lib()
    f = lambda *a: lib.xb_prefill_plan_splits(*a)
    assert f(1, 2048, ctypes.c_int64(2048), 28, 4, 148) == 1          # 112 q tiles x 4 kv heads: enough CTAs
    assert f(1, 16, ctypes.c_int64(8192), 28, 4, 148) == 16           # 4 CTAs: min(148 / 4, 8192 / 512)
    assert f(1, 16, ctypes.c_int64(600), 28, 4, 148) == 1             # short KV: no merge pass
    assert f(8, 18, ctypes.c_int64(4096), 28, 4, 148) == 4
    assert f(1, 1, ctypes.c_int64(100000), 64, 8, 148) == 18          # 8 CTAs -> 148 / 8
    lib.xb_prefill_split_workspace_bytes.restype = ctypes.c_int64
    assert lib.xb_prefill_split_workspace_bytes(4, ctypes.c_int64(16), 28, 128) == 4 * 16 * 28 * 129 * 4
    assert lib.xb_prefill_split_workspace_bytes(1, ctypes.c_int64(16), 28, 128) == 0
