"""Turns ncu reports / launch lists (brought back in gpurun_out/) into the small tracked summaries under profiles/.
  python tools/ncu_summary.py full  <report.ncu-rep> <out.md> [title]     # --set full capture -> key metrics per kernel
  python tools/ncu_summary.py list  <launches.csv>    <out.md> [title]     # gpu__time_duration launch list -> shares
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (legacy hmma)"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor dmma %"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instr (tcgen05/TMA issue)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem / block"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__cycles_active.avg", "sm active cycles"),
    ("sm__cycles_elapsed.avg", "sm elapsed cycles"),
    ("smsp__inst_executed.sum", "warp instructions"),
])


def raw_rows(rep):
    """rep: an .ncu-rep (read through `ncu -i`) or the `--page raw --csv` dump of one made on the GPU box."""
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def full(rep, dst, title):
    hdr, units, rows = raw_rows(rep)
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on)", ""]
    tensor_cols = [i for i, k in enumerate(hdr) if "pipe_tensor" in k and "pct" in k]
    for r in rows:
        name = r[hdr.index("Kernel Name")]
        lines += [f"## {name[:140]}", "", "| metric | value | unit |", "|---|---|---|"]
        for k, label in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"| {label} (`{k}`) | {r[i]} | {units[i]} |")
        for i in tensor_cols:
            if hdr[i] not in KEYS:
                lines.append(f"| `{hdr[i]}` | {r[i]} | {units[i]} |")
        stalls = sorted(((float(r[i].replace(",", "")), k) for i, k in enumerate(hdr)
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")),
                        reverse=True)[:6]
        if stalls:
            lines += ["", "top warp stall reasons (warps stalled per issue-active cycle): " +
                      ", ".join(f"{k.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, k in stalls)]
        rd, wr = r[hdr.index("dram__bytes_read.sum")], r[hdr.index("dram__bytes_write.sum")]
        lines += ["", f"traffic = dram read + write = {rd} {units[hdr.index('dram__bytes_read.sum')]} + {wr} "
                  f"{units[hdr.index('dram__bytes_write.sum')]}", ""]
    open(dst, "w").write("\n".join(lines) + "\n")


def launch_list(path, dst, title):
    text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        name = r["Kernel Name"].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    lines = [f"# {title}", "", f"source: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: "
             "compare SHARES, not absolutes)", "", "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{name}` | {n} | {ns / 1e3:.1f} | {ns / 1e3 / n:.2f} | {100 * ns / total:.1f} % |")
    lines += ["", f"total {total / 1e3:.1f} us over {sum(a[0] for a in agg.values())} launches", ""]
    open(dst, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (full if mode == "full" else launch_list)(src, dst, title)
    print("wrote", dst)
