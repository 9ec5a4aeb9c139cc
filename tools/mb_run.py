"""Run the integer-pipe micro-benchmarks once (the command profiled next to the sumcheck kernel)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import longfellow_zk_b200 as lf
ctx = lf.Context(0)
print(json.dumps({n: ctx.microbench(i) for i, n in enumerate(["imad_wide", "lop3", "gf128_mul", "sha256"])}))
