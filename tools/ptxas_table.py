"""Summarise `ptxas -v` output (XB_PTXAS_V=1 python xllm_b200/build.py -v) as one line per kernel:
demangled name, registers, spill bytes, static smem."""
import re
import subprocess
import sys

txt = sys.stdin.read()
cur = None
rows = []
for line in txt.split("\n"):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = {"name": m.group(1), "spill": 0, "regs": 0, "smem": 0}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        cur["spill"] = int(m.group(1)) + int(m.group(2))
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
        m2 = re.search(r"(\d+) bytes smem", line)
        cur["smem"] = int(m2.group(1)) if m2 else 0
names = subprocess.run(["c++filt"] + [r["name"] for r in rows], capture_output=True, text=True).stdout.split("\n")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for r, n in zip(rows, names):
    short = re.sub(r"\(.*", "", n).replace("void xb::", "")
    if pat in short:
        print(f"{short:70s} regs {r['regs']:3d} spill {r['spill']:3d} smem {r['smem']}")
