"""Static evidence from the shipped library (no GPU needed): per-kernel registers / stack / shared memory from
`cuobjdump --dump-resource-usage` and the count of tensor-core / TMA / async-copy SASS mnemonics per kernel from `cuobjdump -sass`.
  python tools/sass_census.py > profiles/r02_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "xllm_b200", "lib", "libxllm_b200_ops.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "LDGSTS", "HMMA", "QMMA", "SYNCS", "REDG", "ATOMG", "MUFU.EX2"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(d):
    d = re.sub(r"\(.*$", "", d)
    d = d.replace("void ", "")
    return d if len(d) < 96 else d[:93] + "..."


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(v) for v in m.groups())
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in MNEMONICS:
                if op.startswith(k):
                    counts[cur][k] += 1
    names = sorted(usage)
    dm = demangle(names)
    print("# r02 — static census of `libxllm_b200_ops.so` (sm_100a)\n")
    print("`python tools/sass_census.py` at HEAD; registers / stack / static shared memory from `cuobjdump --dump-resource-usage`, SASS\n"
          "mnemonic counts from `cuobjdump -sass` (UTCHMMA / UTCQMMA = tcgen05.mma kind::f16 / kind::f8f6f4, LDTM / STTM = tcgen05.ld / st,\n"
          "UTMALDG / UTMASTG = TMA tensor load / store, LDGSTS = cp.async, HMMA / QMMA = mma.sync bf16 / e4m3, SYNCS = mbarrier ops).\n")
    cols = [k for k in MNEMONICS if any(counts[n][k] for n in names)]
    print("| kernel | regs | stack B | static smem B | SASS instr | " + " | ".join(cols) + " |")
    print("|---|---:|---:|---:|---:|" + "---:|" * len(cols))
    tot = collections.Counter()
    for n in names:
        r = usage[n]
        c = counts[n]
        tot.update(c)
        print(f"| `{short(dm[n])}` | {r[0]} | {r[1]} | {r[2]} | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
    print(f"| **total ({len(names)} kernels)** | | | | {tot['_total']} | " + " | ".join(str(tot[k]) for k in cols) + " |")
    spills = [short(dm[n]) for n in names if usage[n][3] > 0]
    print(f"\nKernels with local-memory spills (LOCAL > 0): {', '.join(spills) if spills else 'none'}.")


if __name__ == "__main__":
    sys.exit(main())
