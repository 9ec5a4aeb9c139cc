"""On the GPU box: condense the source page of an .ncu-rep into the top-N lines by
warp-stall samples (the full report with sources is too large to bring back)."""
import csv
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
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
    print(out[:2000])
    sys.exit(0)
si = next(i for i, c in enumerate(hdr) if "Sampling Data (All)" in c or c.startswith("# Samples"))
srci = next((i for i, c in enumerate(hdr) if c == "Source"), 1)
tot = 0
agg = {}
for r in data:
    try:
        v = float(r[si])
    except Exception:
        continue
    tot += v
    key = r[srci].strip()[:110]
    agg[key] = agg.get(key, 0) + v
print("total samples", tot)
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:n]:
    print(f"{100 * v / max(tot, 1):6.2f}%  {k}")
