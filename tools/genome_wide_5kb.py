#!/usr/bin/env python
"""BASELINE.json configs[3]: genome-wide 5 kb call, every chromosome as a trivial assembly,
given as its band (n x 414) -- SURVEY.md D8.  Chromosomes are sharded over ranks (LPT on length)
when launched under torchrun; otherwise one GPU does all of them.

    python tools/genome_wide_5kb.py [--chroms chr20,chr21,chr22] [--res 5000]

Untimed: procedural map generation and gap removal (host numpy).  Timed: host band -> called loops
(H2D, K5-K10, D2H) for all chromosomes of the rank, per pass.
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neoloopfinder_b200 import _lib
from neoloopfinder_b200.coolshim import HG38_CHROMSIZES, SyntheticCooler
from neoloopfinder_b200.forest import load_forest
from neoloopfinder_b200.scoring import ScoreBatch, score_and_pool
from neoloopfinder_b200.sharding import lpt_shards
from neoloopfinder_b200.synth import analytic_expected

MODEL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "forest_w6.joblib")


def gap_free_band(band_wide, ndiag=414):
    """remove_gaps (callers.py:504-522) on a banded matrix: bins whose column sum is zero go away
    and the band is re-indexed in gap-free coordinates."""
    n, wd = band_wide.shape
    colsum = np.zeros(n)
    for d in range(wd):
        colsum[d:] += band_wide[:n - d, d]
    kept = np.where(colsum != 0)[0]
    m = kept.size
    idx = np.minimum(np.arange(m)[:, None] + np.arange(ndiag)[None, :], m - 1)
    off = kept[idx] - kept[:, None]
    ok = (np.arange(m)[:, None] + np.arange(ndiag)[None, :] < m) & (off < wd)
    out = np.where(ok, band_wide[kept[:, None], np.minimum(off, wd - 1)], 0.0)
    return np.ascontiguousarray(out), kept


class _Only:
    """View of a cooler-like source that lists only some chromosomes but keeps the source's pixel values."""

    def __init__(self, clr, names):
        self._clr = clr
        self.chromnames = list(names)
        self.binsize = clr.binsize
        self.chromsizes = {c: int(clr.chromsizes[c]) for c in names}

    def _raw_block(self, *a):
        return self._clr._raw_block(*a)

    def bins(self):
        return self._clr.bins()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chroms", default=None)
    ap.add_argument("--res", type=int, default=5000)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--prob", type=float, default=0.9)
    ap.add_argument("--pixels", action="store_true",
                    help="row f-2: the map is a pixel table resident in HBM; gap removal and the bands are built on the GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch, torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    names = args.chroms.split(",") if args.chroms else [c for c in HG38_CHROMSIZES if c != "chrY"]
    sizes = {c: HG38_CHROMSIZES[c] for c in names}
    mine = [names[k] for k in lpt_shards([float(sizes[c]) for c in names], world)[rank]]
    clr = SyntheticCooler(binsize=args.res, chromsizes=sizes, seed=404, amplitude=255.0 * args.res / 10000.0)
    exp = analytic_expected(clr)
    exp_arr = np.array([exp[i] for i in sorted(exp)])
    t0 = time.perf_counter()
    bands = []
    ctx = _lib.Context.get(local)
    dp = None
    if args.pixels:
        from neoloopfinder_b200.coolshim import PixelCooler
        sub = SyntheticCooler(binsize=args.res, chromsizes={c: sizes[c] for c in names}, seed=404,
                              amplitude=255.0 * args.res / 10000.0)
        # same procedural pixels as `clr` needs the same global bin offsets: keep all names, fetch only mine
        px = PixelCooler.from_cooler(_Only(sub, mine), band_bins=449, trans=False)
        prep = time.perf_counter() - t0
        t1 = time.perf_counter()
        dp = px.device_pixels(ctx, "sweight")
        ctx.sync()
        upload = time.perf_counter() - t1
        ranges = [px.global_bins(c, 0, px._nbins(c)) for c in mine]
    else:
        for c in mine:
            bw = clr.fetch_band(c, 450, "sweight")
            g, kept = gap_free_band(bw)
            bands.append(g)
        prep = time.perf_counter() - t0
        upload = 0.0
    forest = load_forest(MODEL, ctx)
    times, scored, loops_n = [], 0, 0
    prof = {}
    for p in range(args.passes):
        if p == args.passes - 1:
            ctx.profile(True); ctx.profile_reset()
        ctx.sync(); t0 = time.perf_counter()
        b = ScoreBatch(4, 400, ctx)
        nb = 0
        for g in bands:
            b.add_band(g)
            nb += g.shape[0]
        if dp is not None:
            for lo, hi in ranges:
                _job, kept = b.add_pixels(dp, lo, hi, balance=True)
                nb += kept.size
        loops, (nc, ns, nd) = score_and_pool(b, forest, exp_arr, args.prob, args.res, 2, None, False)
        b.close(); ctx.sync()
        times.append(time.perf_counter() - t0)
        scored = int(ns.sum()); loops_n = sum(len(l[0]) for l in loops)
        if p == args.passes - 1:
            prof = {k: round(ctx.profile_read(v)[0], 2) for k, v in _lib.PROF.items()}
            prof["donuts"] = int(nd.sum())
            ctx.profile(False)
    t = min(times[1:-1]) if len(times) > 2 else min(times)        # the last pass carries profiling events
    if world > 1:
        import torch, torch.distributed as dist
        v = torch.tensor([t], dtype=torch.float64, device="cuda"); dist.all_reduce(v, op=dist.ReduceOp.MAX)
        s = torch.tensor([float(scored), float(loops_n)], dtype=torch.float64, device="cuda"); dist.all_reduce(s)
        t = float(v.item()); scored = int(s[0].item()); loops_n = int(s[1].item())
    if rank == 0:
        print(json.dumps({"workload": "cfg4 genome-wide %d bp, %d chromosomes, banded input" % (args.res, len(names)),
                          "n_gpus": world, "bins": int(nb) if world == 1 else None,
                          "input": "pixel table resident in HBM (gap removal + bands on the GPU)" if args.pixels else "host gap-free bands (H2D per pass)",
                          "pixel_table": ({"nnz": int(px.count.size), "upload_s": round(upload, 4)} if args.pixels else None),
                          "scored_pixels": scored, "loops": loops_n, "seconds_per_pass": t,
                          "pixels_per_s": scored / t, "prep_seconds_rank0": prep, "pass_times": times,
                          "kernel_ms_last_pass_rank0": prof}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
