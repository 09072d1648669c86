"""integration/patches/*.diff are unified diffs against the reference tree (the C++ sides of the
AWQ/GPTQ load path and of the CUDA llama registry entry).  Where the reference checkout is present (the authoring
container) they must apply cleanly; elsewhere only their shape is checked."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCHES = sorted(f for f in os.listdir(os.path.join(ROOT, "integration", "patches")) if f.endswith(".diff"))
REF = "/root/reference"


def _touched(path):
    files = []
    for ln in open(path):
        m = re.match(r"^--- a/(\S+)", ln)
        if m:
            files.append(m.group(1))
    return files


def test_patch_set_is_present():
    assert len(PATCHES) >= 2
    for p in PATCHES:
        body = open(os.path.join(ROOT, "integration", "patches", p)).read()
        assert body.startswith("--- a/xllm/") and "+++ b/xllm/" in body and "@@" in body


@pytest.mark.parametrize("name", PATCHES)
def test_patch_applies_to_the_reference(name, tmp_path):
    if not os.path.isdir(os.path.join(REF, "xllm")):
        pytest.skip("reference checkout not present on this machine")
    path = os.path.join(ROOT, "integration", "patches", name)
    for f in _touched(path):
        src = os.path.join(REF, f)
        if os.path.exists(src):                       # new files have no original
            dst = tmp_path / f
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copy(src, dst)
    r = subprocess.run(["patch", "-p1", "--dry-run", "-i", path], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
