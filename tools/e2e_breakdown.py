"""Where does the end-to-end step go?  (tuning helper, run under gpurun)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from neoloopfinder_b200 import _lib
from neoloopfinder_b200.forest import load_forest
from neoloopfinder_b200.scoring import ScoreBatch, score_and_pool

ctx = _lib.Context.get(0)
forest = load_forest(bench.MODEL, ctx)
WL = bench.Workload(os.environ.get("NLF_TOOL_WORKLOAD", "cfg2"), 50)
clr, lines, expected = WL.build()
exp_arr = np.r_[[expected[i] for i in sorted(expected)]]
mats, chains = bench.gap_free_matrices(clr, list(lines), lines, expected, ctx, forest, WL.res)
pinned = [torch.from_numpy(G).pin_memory() for G in mats]
host = [t.numpy() for t in pinned]
for rep in range(6):
    ctx.sync(); t0 = time.perf_counter()
    b = ScoreBatch(4, 400, ctx)
    t1 = time.perf_counter()
    for G in host:
        b.add_dense(G)
    t2 = time.perf_counter(); ctx.sync(); t3 = time.perf_counter()
    b.score(forest, exp_arr, 0.9, 6)
    t4 = time.perf_counter(); ctx.sync(); t5 = time.perf_counter()
    n = b.counts(); t6 = time.perf_counter()
    don = [b.donuts(j) for j in range(b.njobs)]; t7 = time.perf_counter()
    b.close(); ctx.sync(); t8 = time.perf_counter()
    print("create %.2f add(host) %.2f add(sync) %.2f score(host) %.2f score(sync) %.2f counts %.2f donuts %.2f close %.2f | total %.2f ms"
          % tuple(1e3 * x for x in (t1-t0, t2-t1, t3-t2, t4-t3, t5-t4, t6-t5, t7-t6, t8-t7, t8-t0)))
