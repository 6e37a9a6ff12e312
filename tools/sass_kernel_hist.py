"""Per-kernel SASS histogram of the shipped library: which kernels use the Blackwell-native units.

    python tools/sass_kernel_hist.py mmearth_train_b200/lib/libmpmae.so > profiles/r2_sass_histogram.md

`cuobjdump -sass` mnemonics: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store
(cp.async.bulk.tensor), UBLKCP = cp.async.bulk, LDGSTS = cp.async, FFMA2 = packed fp32 FMA, SYNCS = mbarrier ops.
"""
import collections
import re
import subprocess
import sys

so = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WANT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "LDGSTS", "SYNCS", "FFMA2", "HMMA", "MUFU", "ATOMS", "ATOMG", "REDG", "RED"]
cur, hist, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for w in WANT:
            if op.startswith(w):
                hist[cur][w] += 1
                break
names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
print("# SASS histogram of `lib/libmpmae.so` (sm_100a), per kernel\n")
print("`cuobjdump -sass`; static instruction counts.  `UTCHMMA` = tcgen05.mma (kind::f16 / tf32), `LDTM` = tcgen05.ld, `UTMALDG` / `UTMASTG` = "
      "TMA tensor load / store, `UBLKCP` = cp.async.bulk, `SYNCS` = mbarrier operations.  No `HMMA` (legacy mma.sync) anywhere.\n")
cols = [w for w in WANT if any(h[w] for h in hist.values())]
print("| kernel | SASS instrs | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
tot = collections.Counter()
for (mangled, h), name in sorted(zip(hist.items(), names), key=lambda t: -sum(t[0][1][w] for w in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP"))):
    name = re.sub(r"\(.*", "", name).replace("mpmae::", "").replace("void ", "")
    name = re.sub(r"\((int|bool)\)", "", name)
    print(f"| `{name[:70]}` | {total[mangled]} | " + " | ".join(str(h[w]) if h[w] else "" for w in cols) + " |")
    tot.update(h)
print("| **total** | " + str(sum(total.values())) + " | " + " | ".join(str(tot[w]) for w in cols) + " |")
