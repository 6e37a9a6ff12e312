"""Side-by-side table of the sweeps tools/role_sweep.sh run wrote (one column per library variant)."""
import collections
import re
import sys

rows, tags = collections.OrderedDict(), []
for line in open(sys.argv[1]):
    if line.startswith("== variant"):
        tags.append(line.split()[-1])
        continue
    m = re.match(r"(\S+\s+\S+)\s+M=.*:\s+([\d.]+)", line)
    if m:
        rows.setdefault(m.group(1), []).append(m.group(2))
print("%-16s" % "shape" + "".join("%8s" % t for t in tags))
for k, v in rows.items():
    print("%-16s" % k + "".join("%8s" % x for x in v))
