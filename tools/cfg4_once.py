"""One cfg4 run (N=500, B=8) for profiling: python tools/cfg4_once.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.configs_bench import run
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
r = run("cfg4 N=500 B=%d" % B, B, 500, 0, 2, 2)
print({k: v for k, v in r["kernels_ms_per_step"].items()}, r["ms_per_step"])
