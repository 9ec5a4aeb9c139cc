"""On the GPU box: top SASS instructions of an .ncu-rep by warp-stall samples, with the stall
reasons of each (the full report with sources is too large to bring back)."""
import csv
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if hdr is None:
        if any("Sampling" in c for c in r):
            hdr = r
        continue
    data.append(r)
if hdr is None:
    print(out[:3000])
    sys.exit(0)
print("COLUMNS:", hdr)
si = next(i for i, c in enumerate(hdr) if "Sampling Data (All)" in c or c.startswith("# Samples"))
srci = next((i for i, c in enumerate(hdr) if c == "Source"), 1)
stall_cols = [i for i, c in enumerate(hdr) if c.lower().startswith("stall") or "stall_" in c.lower()]
tot = 0.0
recs = []
for idx, r in enumerate(data):
    try:
        v = float(r[si])
    except Exception:
        continue
    tot += v
    recs.append((v, idx, r))
print("total samples", tot)
# reason totals
agg = {}
for v, idx, r in recs:
    for i in stall_cols:
        try:
            x = float(r[i])
        except Exception:
            continue
        agg[hdr[i]] = agg.get(hdr[i], 0) + x
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:12]:
    print(f"  {k:40s} {100 * v / max(tot, 1):6.2f}%")
for v, idx, r in sorted(recs, key=lambda x: -x[0])[:n]:
    reasons = []
    for i in stall_cols:
        try:
            x = float(r[i])
        except Exception:
            continue
        if x > 0.15 * v:
            reasons.append(f"{hdr[i]}={100 * x / v:.0f}%")
    print(f"{100 * v / max(tot, 1):6.2f}%  #{idx:6d}  {r[srci].strip()[:70]:70s} {' '.join(reasons)}")
