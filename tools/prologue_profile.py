#!/usr/bin/env python
"""calculate_expected from a resident pixel table, for profiling under ncu (gpurun):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:"k_exp|k_px" --csv --log-file gpurun_out/prologue_launches.csv python tools/prologue_profile.py
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neoloopfinder_b200 import _lib
from neoloopfinder_b200.coolshim import PixelCooler, SyntheticCooler
from neoloopfinder_b200.scoring import ScoreBatch
from neoloopfinder_b200.util import calculate_expected

ap = argparse.ArgumentParser()
ap.add_argument("--chrom-mb", type=int, default=248)
ap.add_argument("--res", type=int, default=10000)
args = ap.parse_args()
nd = 5_000_000 // args.res + 1
clr = SyntheticCooler(binsize=args.res, chromsizes={"chr1": args.chrom_mb * 1_000_000}, seed=5,
                      amplitude=255.0 * args.res / 10000.0)
px = PixelCooler.from_cooler(clr, band_bins=nd + 19, trans=False)
ctx = _lib.Context.get(0)
dp = px.device_pixels(ctx, "sweight")
for _ in range(2):
    t0 = time.perf_counter()
    e = calculate_expected(px, ["chr1"], maxdis=5_000_000, balance="sweight")
    print("calculate_expected %.4f s, e[100] = %r" % (time.perf_counter() - t0, e[100]))
    b = ScoreBatch(4, 400, ctx)
    lo, hi = px.global_bins("chr1", 0, px._nbins("chr1"))
    t0 = time.perf_counter()
    _job, kept = b.add_pixels(dp, lo, hi)
    ctx.sync()
    print("add_pixels %.4f s, kept %d" % (time.perf_counter() - t0, kept.size))
    b.close()
