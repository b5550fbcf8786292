"""Timeline of one resident step under the three score modes (tuning helper, run under gpurun):
NLF_TRACE=1 NLF_SCORE_MODE=coop python tools/score_modes.py [n_assemblies]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from neoloopfinder_b200 import _lib
from neoloopfinder_b200.forest import load_forest
from neoloopfinder_b200.scoring import ScoreBatch, score_and_pool
from neoloopfinder_b200.synth import regional_pixel_cooler

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
ctx = _lib.Context.get(0)
forest = load_forest(bench.MODEL, ctx)
WL = bench.Workload(os.environ.get("NLF_TOOL_WORKLOAD", "cfg3"), n)
clr_syn, lines, expected = WL.build()
clr = regional_pixel_cooler(clr_syn, lines, bench.SPAN, nproc=os.cpu_count())
exp_arr = np.r_[[expected[i] for i in sorted(expected)]]
mats, chains = bench.gap_free_matrices(clr, list(lines), lines, expected, ctx, forest, WL.res, nproc=os.cpu_count())
b = ScoreBatch(4, 400, ctx)
for G in mats:
    b.add_dense(G)
ctx.sync()
for rep in range(3):
    b.reset()
    score_and_pool(b, forest, exp_arr, 0.9, WL.res, 2, chains, False)
ctx.sync()
ctx.profile(True)
ctx.profile_reset()
b.reset()
t0 = time.perf_counter()
loops, counts = score_and_pool(b, forest, exp_arr, 0.9, WL.res, 2, chains, False)
ctx.sync()
print("step %.2f ms, scored %d" % (1e3 * (time.perf_counter() - t0), int(counts[1].sum())))
for k, v in _lib.PROF.items():
    print(k, ctx.profile_read(v))
