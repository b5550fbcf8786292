"""Host-side phases of one resident step (tuning helper, run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from neoloopfinder_b200 import _lib, scoring
from neoloopfinder_b200.forest import load_forest
from neoloopfinder_b200.scoring import ScoreBatch

ctx = _lib.Context.get(0)
forest = load_forest(bench.MODEL, ctx)
WL = bench.Workload(os.environ.get("NLF_TOOL_WORKLOAD", "cfg2"), 50)
clr, lines, expected = WL.build()
exp_arr = np.r_[[expected[i] for i in sorted(expected)]]
mats, chains = bench.gap_free_matrices(clr, list(lines), lines, expected, ctx, forest, WL.res)
b = ScoreBatch(4, 400, ctx)
for G in mats:
    b.add_dense(G)
ctx.sync()
P = time.perf_counter
for rep in range(6):
    ctx.sync(); t0 = P()
    ctx.l2_flush(); ctx.sync(); t1 = P()
    b.reset(); ctx.sync(); t2 = P()
    b.score(forest, exp_arr, 0.9, 6); t3 = P()
    ctx.sync(); t4 = P()
    b.counts(); t5 = P()
    b.donuts_all(); t6 = P()
    b._counts = None
    scoring._pool_scored([b], 0.9, 10000, 2, chains, False, 15000); t7 = P()
    print("flush %.3f reset %.3f score(host) %.3f score(gpu) %.3f counts %.3f donuts %.3f pool-all(cached donuts) %.3f | total %.3f ms"
          % tuple(1e3 * x for x in (t1-t0, t2-t1, t3-t2, t4-t3, t5-t4, t6-t5, t7-t6, t7-t0)))
