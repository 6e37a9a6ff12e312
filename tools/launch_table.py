"""Per-launch device times of the LAST step of an ncu launch list (gpu__time_duration.sum CSV), in launch order.

    python tools/launch_table.py gpurun_out/launches.csv [--steps 2]
"""
import argparse
import csv
import re

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
lines = [l for l in open(a.csv) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
idx = [i for i, r in enumerate(rows) if "mask_kernel" in r["Kernel Name"]]
start = idx[-1] if idx else len(rows) - len(rows) // a.steps
tot = 0.0
for r in rows[start:]:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("mpmae::", "").replace("void ", "").strip()
    v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r["Metric Unit"]]
    tot += v
    print(f"{r['ID']:>5} {name[:58]:58s} {v:9.1f} us  grid {r['Grid Size']:>14} block {r['Block Size']}")
print(f"total {tot / 1e3:.3f} ms over {len(rows) - start} launches")
