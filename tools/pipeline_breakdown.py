"""Where does the full pipeline (assembly -> called loops) spend its time?  (tuning helper, gpurun)"""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from neoloopfinder_b200 import _lib
from neoloopfinder_b200.cli import call_assemblies
from neoloopfinder_b200.forest import load_forest

ctx = _lib.Context.get(0)
forest = load_forest(bench.MODEL, ctx)
WL = bench.Workload(os.environ.get("NLF_TOOL_WORKLOAD", "cfg2"), 16)
clr, lines, expected = WL.build()
ids = list(lines)
call_assemblies(clr, ids[:2], lines, bench.SPAN, "sweight", 0.9, 2, False, forest, expected, ctx)   # warm-up
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
out = call_assemblies(clr, ids, lines, bench.SPAN, "sweight", 0.9, 2, False, forest, expected, ctx)
pr.disable()
print("16 assemblies: %.2f s, loops %d" % (time.perf_counter() - t0, sum(len(v) for v in out.values() if v)))
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
for nproc in (1, 8, 16):
    t0 = time.perf_counter()
    out2 = call_assemblies(clr, ids, lines, bench.SPAN, "sweight", 0.9, 2, False, forest, expected, ctx, nproc=nproc)
    print("nproc=%d: %.2f s  same=%s" % (nproc, time.perf_counter() - t0, out2 == out))
