"""Static checks of the product tree: the oracle is test infrastructure (only tests/, __graft_entry__.smoke() and bench.py's CPU
legs may import it), the hot path uses no compatibility layer (Triton / tilelang / torch.compile), nothing that runs on the GPU box
reads /root/reference, and the library loader has no CPU fallback."""
import ast
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(d):
    for base, _, files in os.walk(os.path.join(ROOT, d)):
        if "__pycache__" in base:
            continue
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(base, f)


def _imports(path):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name, node.lineno
        elif isinstance(node, ast.ImportFrom) and node.module:
            yield node.module, node.lineno


def test_product_package_never_imports_the_oracle_or_a_compat_layer():
    bad = []
    for f in _py_files("xllm_b200"):
        for mod, line in _imports(f):
            top = mod.split(".")[0]
            if top in ("oracle", "triton", "tilelang", "tests"):
                bad.append(f"{os.path.relpath(f, ROOT)}:{line} imports {mod}")
        src = open(f).read()
        if re.search(r"torch\.compile\s*\(", src):
            bad.append(f"{os.path.relpath(f, ROOT)} calls torch.compile")
    assert not bad, bad


def test_bench_and_entry_import_the_oracle_only_in_the_cpu_legs_and_smoke():
    # build() may import oracle.build_ref and nothing else of the oracle: it BUILDS the checker (oracle/_ref), it does not use it
    allowed = {"bench.py": {"cpu_layer_baseline"}, "__graft_entry__.py": {"smoke", "build"}}
    for fname, funcs in allowed.items():
        tree = ast.parse(open(os.path.join(ROOT, fname)).read())
        for node in tree.body:                                            # module level: no oracle / tests imports
            if isinstance(node, (ast.Import, ast.ImportFrom)):
                mods = [a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""]
                assert not any(m.split(".")[0] in ("oracle", "tests") for m in mods), f"{fname}:{node.lineno}"
        for node in ast.walk(tree):
            if isinstance(node, ast.FunctionDef):
                uses = [n for n in ast.walk(node) if isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] in ("oracle", "tests")]
                if uses:
                    assert node.name in funcs, f"{fname}: {node.name}() imports the oracle / tests"
                    if node.name == "build":
                        names = {(n.module, a.name) for n in uses for a in n.names}
                        assert names == {("oracle", "build_ref")}, names


def test_nothing_outside_tools_and_fixture_scripts_reads_the_reference_tree():
    offenders = []
    for d in ("xllm_b200", "oracle"):
        for f in _py_files(d):
            for i, line in enumerate(open(f), 1):
                code = line.split("#", 1)[0]
                if "/root/reference" in code and "open(" in code:
                    offenders.append(f"{os.path.relpath(f, ROOT)}:{i}")
    for fname in ("bench.py", "__graft_entry__.py"):
        for i, line in enumerate(open(os.path.join(ROOT, fname)), 1):
            code = line.split("#", 1)[0]
            if "/root/reference" in code and ("open(" in code or "listdir" in code or "import" in code):
                offenders.append(f"{fname}:{i}")
    assert not offenders, offenders


def test_library_loader_raises_when_the_extension_is_missing(tmp_path, monkeypatch):
    import pytest
    from xllm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libxllm_b200_ops.so"))
    with pytest.raises(_lib.XllmB200Error, match="is missing"):
        _lib.lib()
    with pytest.raises(_lib.XllmB200Error):                              # an op call goes through the same loader: no fallback path
        from xllm_b200 import ops
        import torch
        ops.set_w4_decode_form(0)
