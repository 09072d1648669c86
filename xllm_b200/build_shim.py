"""Compiles the link-time drop-in for `xllm::kernel::cuda::*` (csrc/shim/xllm_cuda_ops.cpp) against libtorch and
libxllm_b200_ops.so.  A maintainer would compile that file inside xLLM's own build instead; building it
here proves the signatures of cuda_ops_api.h:31-266 are met and gives tests a library to inspect."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "shim", "xllm_cuda_ops.cpp")
LIBDIR = os.path.join(HERE, "lib")
OUT = os.path.join(LIBDIR, "libxllm_b200_shim.so")


def build(verbose=False, force=False):
    deps = [SRC, os.path.join(HERE, "..", "include", "xllm_b200_ops.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) > os.path.getmtime(d) for d in deps):
        return OUT
    import torch
    ti = os.path.dirname(torch.__file__)
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", SRC, "-o", OUT, f"-I{ti}/include",
           f"-I{ti}/include/torch/csrc/api/include", "-I/usr/local/cuda/include", "-D_GLIBCXX_USE_CXX11_ABI=1", f"-L{LIBDIR}",
           "-lxllm_b200_ops", f"-L{ti}/lib", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda",
           f"-Wl,-rpath,{ti}/lib", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("shim build failed")
    return OUT


PY_SRC = os.path.join(HERE, "csrc", "shim", "shim_py.cpp")
PY_OUT = os.path.join(LIBDIR, "xllm_b200_shim_py.so")


def build_py(verbose=False, force=False):
    """test binding of the shim (csrc/shim/shim_py.cpp): a CPython module exposing xllm::kernel::cuda::* with torch tensors."""
    so = build(verbose, force)
    if not force and os.path.exists(PY_OUT) and all(os.path.getmtime(PY_OUT) > os.path.getmtime(d) for d in (PY_SRC, so)):
        return PY_OUT
    import sysconfig
    import pybind11
    import torch
    ti = os.path.dirname(torch.__file__)
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", PY_SRC, "-o", PY_OUT, f"-I{ti}/include",
           f"-I{ti}/include/torch/csrc/api/include", "-I/usr/local/cuda/include", f"-I{pybind11.get_include()}",
           f"-I{sysconfig.get_paths()['include']}", "-D_GLIBCXX_USE_CXX11_ABI=1", "-DTORCH_EXTENSION_NAME=xllm_b200_shim_py",
           f"-L{LIBDIR}", "-lxllm_b200_shim", "-lxllm_b200_ops", f"-L{ti}/lib", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10",
           f"-Wl,-rpath,{ti}/lib", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("shim python binding build failed")
    return PY_OUT


def load_py():
    """import the test binding (builds it if needed)"""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    so = build_py()
    spec = importlib.util.spec_from_file_location("xllm_b200_shim_py", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
