"""Per-step wall times of the resident and end-to-end arms (tuning helper, run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from neoloopfinder_b200 import _lib
from neoloopfinder_b200.forest import load_forest
from neoloopfinder_b200.scoring import ScoreBatch, score_and_pool, score_matrices

ctx = _lib.Context.get(0)
forest = load_forest(bench.MODEL, ctx)
WL = bench.Workload(os.environ.get("NLF_TOOL_WORKLOAD", "cfg2"), 50)
clr, lines, expected = WL.build()
exp_arr = np.r_[[expected[i] for i in sorted(expected)]]
mats, chains = bench.gap_free_matrices(clr, list(lines), lines, expected, ctx, forest, WL.res)
pinned = [torch.from_numpy(G).pin_memory() for G in mats]
host = [t.numpy() for t in pinned]
resident = ScoreBatch(4, 400, ctx)
for G in host:
    resident.add_dense(G)
ctx.sync()
for rep in range(5):
    t0 = time.perf_counter(); ctx.l2_flush(); resident.reset()
    score_and_pool(resident, forest, exp_arr, 0.9, 10000, 2, chains, False); ctx.sync()
    print("resident step %.2f ms" % (1e3 * (time.perf_counter() - t0)))
for rep in range(8):
    t0 = time.perf_counter(); ctx.l2_flush()
    score_matrices(host, forest, exp_arr, 0.9, 10000, 2, chains, False, 400, ctx); ctx.sync()
    print("e2e step %.2f ms" % (1e3 * (time.perf_counter() - t0)))
