"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
usage: python tools/launch_agg.py gpurun_out/x.csv [steps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    n = r[kn].split("(")[0][:70]
    agg[n][0] += 1
    agg[n][1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print("per step (%d steps in the capture); ncu serialises launches and runs them cold" % steps)
for n, v in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print(f"{n:72s} {v[0] / steps:7.1f} launches {v[1] / 1e6 / steps:9.3f} ms {100 * v[1] / tot:5.1f}%")
print("total %.3f ms per step" % (tot / 1e6 / steps))
