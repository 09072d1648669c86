"""CPU checks of the drop-in boundary: the library loads and exports every symbol include/xllm_b200_ops.h declares;
host-side packers agree; argument errors are reported through the C ABI without touching a GPU."""
import ctypes
import os
import re

import torch

from xllm_b200 import _lib, quant

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "xllm_b200_ops.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert _lib.lib().xb_abi_version() == 1


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "xllm_b200_ops.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert 'extern "C"' in src


def test_w4_packers_agree(built_lib):
    g = torch.Generator().manual_seed(2026)
    q = torch.randint(0, 16, (64, 256), dtype=torch.uint8, generator=g)
    s = torch.rand(64, 2, generator=g).to(torch.bfloat16)
    z = torch.randint(0, 16, (64, 2), dtype=torch.uint8, generator=g)
    qa, meta = quant.pack_w4(q, s, z, 128)
    assert torch.equal(qa, quant.pack_w4_c(q))
    # meta = bf16(scale) | bf16(128+zero) << 16, laid out [K/g, N]
    assert meta.shape == (2, 64)
    lo = (meta & 0xFFFF).to(torch.int16).view(torch.bfloat16)
    hi = ((meta >> 16) & 0xFFFF).to(torch.int16).view(torch.bfloat16)
    assert torch.equal(lo.t().contiguous(), s)
    assert torch.equal(hi.t().to(torch.float32), z.to(torch.float32) + 128.0)


def test_w8_packers_agree_and_qkv_rope_index(built_lib):
    g = torch.Generator().manual_seed(2026)
    q = torch.randint(0, 256, (64, 256), dtype=torch.uint8, generator=g)
    s = torch.rand(64, 2, generator=g).to(torch.bfloat16)
    z = torch.randint(0, 256, (64, 2), dtype=torch.uint8, generator=g)
    qa, meta = quant.pack_w8(q, s, z, 128)
    assert qa.shape == (4, 4, 32, 8) and torch.equal(qa, quant.pack_w8_c(q))
    # lane 4g+t of tile (nt, kt): words 0..3 = row 16nt+g, bytes k = 64kt+16t .. +15 ascending; words 4..7 = row +8
    raw = qa.view(torch.uint8).view(4, 4, 32, 2, 16)
    assert torch.equal(raw[1, 2, 4 * 3 + 1, 0], q[16 + 3, 128 + 16:128 + 32]) and torch.equal(raw[1, 2, 4 * 3 + 1, 1], q[16 + 11, 128 + 16:128 + 32])
    assert torch.equal((meta & 0xFFFF).to(torch.int16).view(torch.bfloat16).t().contiguous(), s)
    assert torch.equal((meta >> 16).t().contiguous().to(torch.uint8), z)
    # rope-pair row order of the decode qkv layout: inside each head, tile j = dims 8j..8j+7 then D/2+8j..D/2+8j+7
    idx = quant.qkv_rope_index(2, 1, 64)
    assert sorted(idx.tolist()) == list(range(4 * 64))
    assert idx[:16].tolist() == list(range(8)) + list(range(32, 40)) and idx[64 + 16:64 + 24].tolist() == list(range(64 + 8, 64 + 16))


def test_argument_errors_surface(built_lib):
    lib = _lib.lib()
    plan = (ctypes.c_int64 * 8)()
    rc = lib.xb_decode_plan(plan, 1, 28, 4, 96, 128, 32, 148)      # head_dim 96 unsupported
    assert rc != 0 and b"head_dim" in lib.xb_last_error()
    rc = lib.xb_decode_plan(plan, 1, 28, 4, 128, 128, 32, 148)
    assert rc == 0 and plan[0] % 16 == 0 and plan[0] * plan[1] >= 4096
    rc = lib.xb_w4_pack_rows(None, None, 15, 64)
    assert rc != 0
    assert lib.xb_w8_pack_rows(None, None, 16, 60) != 0
    # the fused decode GEMV serves one token tile and needs its activation block to fit the shared-memory stage
    assert lib.xb_linear_w4a16_decode_fused_fits(8, 3584) == 1 and lib.xb_linear_w4a16_decode_fused_fits(9, 3584) == 0
    assert lib.xb_linear_w4a16_decode_fused_fits(1, 18944) == 1 and lib.xb_linear_w4a16_decode_fused_fits(8, 18944) == 0
    assert lib.xb_linear_w8a16_small_m(None, 0, None, 0, None, None, None, 65, 64, 64, 64, None) != 0
    assert b"M=65" in lib.xb_last_error()
    assert lib.xb_linear_fp8_small_m(None, 0, None, 0, None, None, 1, None, 1, None, 4, 64, 96, None) != 0


def test_tvm_ffi_modules_export_reference_entry_points(built_lib):
    """xLLM resolves __tvm_ffi_<name> in "$FLASHINFER_OPS_PATH/<uri>/<uri>.so" (utils.cpp:371-374,526-564)."""
    from xllm_b200 import build_ffi
    ops_dir = build_ffi.build()
    dec, pre = build_ffi.uris()
    for uri in dec:
        so = ctypes.CDLL(os.path.join(ops_dir, uri, uri + ".so"))
        assert hasattr(so, "__tvm_ffi_plan") and hasattr(so, "__tvm_ffi_run")
        assert hasattr(so, "__tvm_ffi_plan_is_replay_invariant")     # capability probe of integration/patches/0003
    for uri in pre:
        so = ctypes.CDLL(os.path.join(ops_dir, uri, uri + ".so"))
        assert all(hasattr(so, "__tvm_ffi_" + n) for n in ("plan", "ragged_run", "paged_run"))


def test_cpp_shim_exports_reference_namespace(built_lib):
    """the link-time boundary: every xllm::kernel::cuda::* function of cuda_ops_api.h that is on the path."""
    import subprocess
    from xllm_b200 import build_shim
    so = build_shim.build()
    syms = subprocess.run(["nm", "-D", "-C", so], capture_output=True, text=True).stdout
    for fn in ("rotary_embedding", "act_and_mul", "reshape_paged_cache", "rms_norm", "fused_add_rms_norm", "matmul",
               "cutlass_scaled_mm", "static_scaled_fp8_quant", "fp8_scaled_quantize", "rms_norm_static_fp8_quant",
               "fused_add_rms_norm_static_fp8_quant", "fp8_scaled_matmul", "fused_qk_norm_rope",
               "update_llm_decode_metadata", "moe_fused_topk", "cutlass_fused_moe"):
        assert f"xllm::kernel::cuda::{fn}(" in syms, fn


def test_header_is_plain_c_and_cxx(tmp_path):
    """the boundary a cgo / JNI / ctypes / C++ caller binds: include/xllm_b200_ops.h must compile as C11 (pedantic) and as C++17 on
    its own - no C++-only constructs, no missing includes"""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "xllm_b200_ops.h"\nint main(void) { return 0; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)],
                ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c++", str(src)]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_plain_c_caller_links_and_runs(tmp_path, built_lib):
    """a C program - what a cgo / JNI / FFI binding boils down to - includes the header, links the shared library and calls
    host-only entry points (no device needed): ABI version, the GEMM dispatch description, a decode plan, an argument error with its
    message."""
    import subprocess
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "xllm_b200_ops.h"
int main(void) {
  char buf[96];
  long long plan[8];
  int64_t p64[8];
  if (xb_abi_version() != 1) return 1;
  if (xb_gemm_describe(1, 32, 1280, 8192, 148, buf, (int)sizeof buf) <= 0) return 2;
  printf("%s\n", buf);
  if (xb_decode_plan(p64, 3, 28, 4, 128, 16, 300, 148) != 0) return 3;
  for (int i = 0; i < 8; ++i) plan[i] = (long long)p64[i];
  printf("plan %lld %lld\n", plan[4], plan[5]);
  if (xb_decode_plan(p64, 3, 28, 5, 128, 16, 300, 148) == 0) return 4;      /* 28 q heads over 5 kv heads: rejected */
  printf("err %s\n", strlen(xb_last_error()) > 0 ? "set" : "empty");
  return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(built_lib)
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        f"-L{libdir}", "-lxllm_b200_ops", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    out = r.stdout.splitlines()
    assert out[0] == "fp8 swap-AB bn=32 split_k=5" and out[1] == "plan 3 28" and out[2] == "err set"


def test_every_abi_entry_states_what_it_replaces():
    """include/*.h declares the boundary "citing the reference interface each one replaces (file:line)": the comment in front of every
    xb_* declaration either cites a reference source line or says that the entry is additive / library-level / a host-only query / a
    tuning setter (things the reference has no counterpart for)."""
    src = open(os.path.join(ROOT, "include", "xllm_b200_ops.h")).read()
    last, missing = "", []
    for m in re.finditer(r"/\*.*?\*/|\b(?:int|size_t|int64_t|const char\*|uint64_t|void)\s+(xb_[a-z0-9_]+)\s*\(", src, flags=re.S):
        if m.group(0).startswith("/*"):
            last = m.group(0)
        elif not re.search(r"\.(cpp|cu|cuh|h):\d+|additive|library-level|debug aid|host-only|SURVEY|setter|returns the old|Returns the old",
                           last):
            missing.append(m.group(1))
    assert not missing, missing
