"""Summarises an ncu launch list (gpu__time_duration.sum CSV) and optionally a --set full report into Markdown.

    python tools/summarize_ncu.py gpurun_out/launches.csv [--rep gpurun_out/prof.ncu-rep] [--steps 2] > profiles/NAME.md
"""
import argparse
import collections
import csv
import re
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--rep", default=None)
ap.add_argument("--steps", type=int, default=2, help="steps in the capture; the LAST one is summarised")
ap.add_argument("--title", default="ncu launch list")
a = ap.parse_args()

lines = [l for l in open(a.csv) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per = len(rows) // a.steps
rows = rows[len(rows) - per:]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("mpmae::", "").replace("void ", "").strip()
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r["Metric Unit"]]
    e = agg.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {a.title}\n")
print(f"`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES). "
      f"Last of {a.steps} captured steps: {per} launches, {tot:.3f} ms summed kernel time.\n")
print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |")
if a.rep:
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    hdr = rr[0]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    print(f"\n## `ncu --set full` capture: `{a.rep.split('/')[-1]}`\n")
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rr) - 2)) + " |")
    print("|---|---|" + "---:|" * (len(rr) - 2))
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            vals = [re.sub(r"\(.*", "", r[i])[:60] for r in rr[2:]]
            print(f"| `{w}` | {rr[1][i]} | " + " | ".join(vals) + " |")
