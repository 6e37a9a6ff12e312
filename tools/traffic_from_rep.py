"""DRAM traffic per launch of the bench's dominant kernel from an `ncu --set full` report of its launches:
    python tools/traffic_from_rep.py gpurun_out/X_full_pw1.ncu-rep pw1 "profiles/X_launches.md" > profiles/traffic.json"""
import csv
import json
import subprocess
import sys

rep, name, src = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines()))
hdr = rr[0]
ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = sum(float(r[ir]) * scale[rr[1][ir]] + float(r[iw]) * scale[rr[1][iw]] for r in rr[2:])
n = len(rr) - 2
json.dump({name: {"traffic_bytes_per_launch": tot / n, "launches": n,
                  "source": f"{src} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum averaged over the {n} {name} "
                            "launches of one step; write-back still resident in the 126 MB L2 at kernel end is not counted)"}},
          sys.stdout, indent=1)
