"""mdoc throughput, two batches of B in flight (tools/mdoc_bench.measure_two_in_flight), for knob sweeps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import mdoc_bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
r = mdoc_bench.measure_two_in_flight(B=B, rounds=2)
print({k: v for k, v in os.environ.items() if k.startswith("LF_")}, "B", B, "proofs/s %.1f" % r["proofs_per_s"])
