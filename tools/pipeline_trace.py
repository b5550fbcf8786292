"""Host-side timeline of the pipelined end-to-end call (tuning helper, run under gpurun):
when each add / score / result phase of every sub-batch returns, for several pipeline depths."""
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
for depth, first in ((1, 1.0), (2, 1.0), (4, 1.0), (4, 0.5), (4, 0.3), (5, 0.5), (5, 0.3), (6, 0.4)):
    os.environ["NLF_PIPELINE_FIRST"] = str(first)
    ts = []
    for rep in range(8):
        ctx.l2_flush(); ctx.sync()
        t0 = time.perf_counter()
        score_matrices(host, forest, exp_arr, 0.9, 10000, 2, chains, False, 400, ctx, depth=depth)
        ctx.sync()
        ts.append(1e3 * (time.perf_counter() - t0))
    print("depth %d first %.1f: e2e step median %.2f ms (min %.2f)" % (depth, first, sorted(ts[2:])[3], min(ts[2:])))
