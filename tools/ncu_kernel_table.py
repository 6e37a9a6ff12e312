"""Per-kernel evidence table from an ncu report of one whole step:

    python tools/ncu_kernel_table.py gpurun_out/step.ncu-rep [--peak-gbs 6538.6] > profiles/NAME.md

For every kernel name: launches, summed duration, DRAM bytes and achieved GB/s (share of the measured copy peak), the
duration-weighted means of ncu's DRAM / SM throughput percentages, tensor-pipe activity, achieved occupancy and IPC."""
import argparse
import collections
import csv
import json
import os
import re
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--peak-gbs", type=float, default=None)
ap.add_argument("--title", default="per-kernel ncu table")
a = ap.parse_args()
peak = a.peak_gbs
if peak is None:
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
out = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines()))
hdr, units, rows = rr[0], rr[1], rr[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, default=0.0):
    if name not in col:
        return default
    try:
        v = float(r[col[name]].replace(",", ""))
    except ValueError:
        return default
    u = units[col[name]].split("/")[0]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return v * scale


agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("mpmae::", "").replace("void ", "").strip()
    t = val(r, "gpu__time_duration.sum")
    e = agg.setdefault(name, dict(n=0, t=0.0, rd=0.0, wr=0.0, dram=0.0, sm=0.0, tc=0.0, occ=0.0, ipc=0.0, regs=0, smem=0.0))
    e["n"] += 1
    e["t"] += t
    e["rd"] += val(r, "dram__bytes_read.sum")
    e["wr"] += val(r, "dram__bytes_write.sum")
    e["dram"] += t * val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    e["sm"] += t * val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
    e["tc"] += t * val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    e["occ"] += t * val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    e["ipc"] += t * val(r, "sm__inst_executed.avg.per_cycle_active")
    e["regs"] = max(e["regs"], int(val(r, "launch__registers_per_thread")))
    e["smem"] = max(e["smem"], val(r, "launch__shared_mem_per_block_dynamic") / 1e3)
tot = sum(e["t"] for e in agg.values())
print(f"# {a.title}\n")
print(f"`ncu` replay of every launch of one step ({len(rows)} launches, {tot / 1e3:.3f} ms summed; cold caches, serialised, so "
      f"durations are upper bounds and SHARES are what compares with the CUDA-event profile of `bench.py`).  DRAM GB/s = "
      f"(dram__bytes_read.sum + dram__bytes_write.sum) / duration; `% of copy peak` is against the measured {peak:.0f} GB/s "
      f"(`MEASURED_PEAKS.json`).  Write-back that is still in the 126 MB L2 when a kernel ends is not counted by the DRAM "
      f"counters.\n")
have_bytes = "dram__bytes_read.sum" in col
bcols = " DRAM MB | GB/s | % of copy peak |" if have_bytes else ""
print(f"| kernel | launches | ms | share |{bcols} ncu DRAM % | SM % | tensor pipe % | warps active % | IPC | regs | smem KB |")
print("|---|---:|---:|---:|" + ("---:|---:|---:|" if have_bytes else "") + "---:|---:|---:|---:|---:|---:|---:|")
for k, e in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    t = e["t"]
    if t <= 0:
        continue
    gbs = (e["rd"] + e["wr"]) / (t * 1e-6) / 1e9
    b = f" {(e['rd'] + e['wr']) / 1e6:.0f} | {gbs:.0f} | {100 * gbs / peak:.1f}% |" if have_bytes else ""
    print(f"| `{k}` | {e['n']} | {t / 1e3:.3f} | {100 * t / tot:.1f}% |{b} {e['dram'] / t:.1f} | {e['sm'] / t:.1f} | "
          f"{e['tc'] / t:.1f} | {e['occ'] / t:.1f} | {e['ipc'] / t:.2f} | {e['regs']} | {e['smem']:.0f} |")
