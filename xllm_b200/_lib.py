"""ctypes binding of libxllm_b200_ops.so (the C ABI in include/xllm_b200_ops.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libxllm_b200_ops.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "xllm_b200_ops.h")

_lib = None


class XllmB200Error(RuntimeError):
    pass


def lib():
    """Load the library (fails loudly when it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise XllmB200Error(
                f"{LIB_PATH} is missing: build it with `python -m xllm_b200.build` "
                "(or __graft_entry__.build()).  There is no CPU/PyTorch fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.xb_last_error.restype = ctypes.c_char_p
        _lib.xb_launch_count.restype = ctypes.c_uint64
        _lib.xb_abi_version.restype = ctypes.c_int
        _lib.xb_prefill_split_workspace_bytes.restype = ctypes.c_int64
        _lib.xb_moe_experts_workspace_bytes.restype = ctypes.c_int64
    return _lib


def check(rc, what):
    if rc != 0:
        raise XllmB200Error(f"{what}: {lib().xb_last_error().decode()}")


def launch_count() -> int:
    return int(lib().xb_launch_count())


def set_pdl(enable: bool):
    lib().xb_set_pdl(ctypes.c_int(1 if enable else 0))


c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int
c_f32 = ctypes.c_float
c_ptr = ctypes.c_void_p
