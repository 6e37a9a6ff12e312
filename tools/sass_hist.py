"""Opcode histogram (executed instructions + stall samples) from an `ncu --page source --csv --print-source sass` export."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ie]) for r in data)
ts = sum(int(r[isamp]) for r in data)
print("total warp-inst", tot, "samples", ts, "static instrs", len(data))
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    t = r[ia].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    ops[op] += int(r[ie])
    samp[op] += int(r[isamp])
for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:12s} {c:10d} {100 * c / tot:5.1f}%  samples {100 * samp[op] / ts:5.1f}%")
if len(sys.argv) > 3:   # top sampled instructions
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:int(sys.argv[3])]:
        print(r[isamp], r[ie], r[ia][:100])
