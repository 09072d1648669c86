"""SASS inspection helpers used during round 1 (no GPU needed; reads the in-tree .so with cuobjdump).

  python tools/sass_inspect.py kernels [substr]                 # list kernel symbols (mangled) matching substr
  python tools/sass_inspect.py hist <mangled-or-substr> [lo hi]  # opcode histogram of the kernel (or of [lo, hi] hex range)
  python tools/sass_inspect.py sb <mangled-or-substr> [lo hi] [regex]
        # per instruction: stall count, write/read scoreboard, wait mask (decoded from the control bits) - this is how
        # the "all LDGs of the ring share scoreboard 5" problem of the register-ring GEMV was found

Control word layout (Volta+, upper 64 bits of the 128-bit instruction): stall[41:44] yield[45] wr_sb[46:48] rd_sb[49:51]
wait_mask[52:57] reuse[58:61].
"""
import collections
import re
import subprocess
import sys

LIB = "xllm_b200/lib/libxllm_b200_ops.so"
PAT = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/")
PAT2 = re.compile(r"/\* 0x([0-9a-f]{16}) \*/")


def kernels(sub=""):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    return [k for k in re.findall(r"Function : (\S+)", out) if sub in k]


def sass(fun):
    names = kernels(fun)
    if not names:
        sys.exit(f"no kernel matching {fun}")
    name = names[0] if fun not in names else fun
    out = subprocess.run(["cuobjdump", "-sass", "-fun", name, LIB], capture_output=True, text=True).stdout.split("\n")
    rows, i = [], 0
    while i < len(out):
        m = PAT.search(out[i])
        if m and i + 1 < len(out):
            h = int(PAT2.search(out[i + 1]).group(1), 16)
            rows.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=(h >> 41) & 0xF, wr=(h >> 46) & 7,
                             rd=(h >> 49) & 7, wait=(h >> 52) & 0x3F))
            i += 2
        else:
            i += 1
    return name, rows


def opcode(text):
    parts = text.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    return op.split(".")[0]


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "kernels":
        print("\n".join(kernels(sys.argv[2] if len(sys.argv) > 2 else "")))
        sys.exit(0)
    name, rows = sass(sys.argv[2])
    lo = int(sys.argv[3], 16) if len(sys.argv) > 4 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
    rows = [r for r in rows if lo <= r["addr"] <= hi]
    print(name, f"({len(rows)} instructions)")
    if cmd == "hist":
        for op, n in collections.Counter(opcode(r["text"]) for r in rows).most_common():
            print(f"{n:6d} {op}")
    elif cmd == "sb":
        rx = re.compile(sys.argv[5]) if len(sys.argv) > 5 else None
        for r in rows:
            if rx is None or rx.search(r["text"]) or r["wait"]:
                wr = r["wr"] if r["wr"] != 7 else "-"
                rd = r["rd"] if r["rd"] != 7 else "-"
                print(f"{r['addr']:05x} st{r['stall']:2d} wr{wr} rd{rd} wait{r['wait']:06b}  {r['text'][:100]}")
