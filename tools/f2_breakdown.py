#!/usr/bin/env python
"""Row f-2 (resident pixel table) and the normalisation kernels, timed on one GPU (gpurun):

  * per-assembly preparation, host path (cooler-surface fetches -> dense matrix -> upload) vs the
    device path (matrix built from the resident pixels), then diag means + gap removal;
  * calculate_expected, host band path vs resident pixels;
  * chromosomes as trivial assemblies: host gap removal + band upload vs add_pixels.

    python tools/f2_breakdown.py [--assemblies 24] [--chrom-mb 40] > gpurun_out/f2_breakdown.json
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neoloopfinder_b200 import _lib
from neoloopfinder_b200.assembly import complexSV
from neoloopfinder_b200.coolshim import PixelCooler, SyntheticCooler
from neoloopfinder_b200.scoring import ScoreBatch
from neoloopfinder_b200.synth import make_workload
from neoloopfinder_b200.util import calculate_expected


def prof(ctx, which="norm"):
    ms, n = ctx.profile_read(_lib.PROF[which])
    return round(ms, 4), int(n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--assemblies", type=int, default=24)
    ap.add_argument("--chrom-mb", type=int, default=40)
    ap.add_argument("--span", type=int, default=3_000_000)
    ap.add_argument("--chrom-res", type=int, default=10000)
    ap.add_argument("--skip-host-band", action="store_true")
    args = ap.parse_args()
    out = {}
    ctx = _lib.Context.get(0)
    sizes = {"chr1": 30_000_000, "chr2": 24_000_000, "chr3": 20_000_000}
    clr, lines = make_workload("simple", res=10000, seed=11, n_assemblies=args.assemblies, chromsizes=sizes, span=args.span)
    t0 = time.perf_counter()
    px = PixelCooler.from_cooler(clr)
    out["pixel_table"] = {"bins": int(px._total_bins), "nnz": int(px.count.size),
                          "host_build_s": round(time.perf_counter() - t0, 2)}
    t0 = time.perf_counter()
    px.device_pixels(ctx, "sweight")
    ctx.sync()
    out["pixel_table"]["upload_s"] = round(time.perf_counter() - t0, 4)
    ctx.profile(True)
    for name, src in (("host", clr), ("pixels", px)):
        # warm-up on one assembly, then time all of them stage by stage
        rec = {}
        for rep in range(2):
            ids = list(lines)[:1] if rep == 0 else list(lines)
            ctx.sync(); ctx.profile_reset()
            t = {"reorganize": 0.0, "diag_means": 0.0, "remove_gaps": 0.0}
            nbins = 0
            for ID in ids:
                fu = complexSV(src, lines[ID], span=args.span, col="sweight", expected_values={}, flexible=False)
                fu._ctx = ctx
                a = time.perf_counter()
                fu.reorganize_matrix()
                asm = fu._device_assembly()
                ctx.sync()
                b = time.perf_counter()
                asm.diag_means(200)
                c = time.perf_counter()
                asm.set_scales(np.zeros((asm.nblk, asm.nblk), np.int32), np.ones((asm.nblk, asm.nblk)))
                asm.kept_index()
                d = time.perf_counter()
                t["reorganize"] += b - a; t["diag_means"] += c - b; t["remove_gaps"] += d - c
                nbins += asm.n
                asm.close()
            ms, n = prof(ctx)
            rec = {"assemblies": len(ids), "mean_bins": nbins / len(ids),
                   "wall_ms_per_assembly": {k: round(1e3 * v / len(ids), 3) for k, v in t.items()},
                   "norm_kernel_ms_per_assembly": round(ms / len(ids), 4), "norm_launches_per_assembly": n / len(ids)}
        out["assembly_" + name] = rec
    # calculate_expected: one long chromosome
    big = {"chr1": args.chrom_mb * 1_000_000}
    cres = args.chrom_res
    nd = 5_000_000 // cres + 1
    c2 = SyntheticCooler(binsize=cres, chromsizes=big, seed=5, amplitude=255.0 * cres / 10000.0)
    t0 = time.perf_counter()
    p2 = PixelCooler.from_cooler(c2, band_bins=nd + 19, trans=False)
    build_s = time.perf_counter() - t0
    p2.device_pixels(ctx, "sweight"); ctx.sync()
    res = {}
    for name, src in ((("pixels", p2),) if args.skip_host_band else (("host_band", c2), ("pixels", p2))):
        calculate_expected(src, ["chr1"], maxdis=5_000_000, balance="sweight")       # warm-up
        ctx.sync(); ctx.profile_reset()
        t0 = time.perf_counter()
        e = calculate_expected(src, ["chr1"], maxdis=5_000_000, balance="sweight")
        wall = time.perf_counter() - t0
        ms, n = prof(ctx)
        res[name] = {"wall_s": round(wall, 4), "kernel_ms": ms, "launches": n}
        res.setdefault("check", []).append(float(e[100]))
    n = p2._total_bins
    res["bins"] = int(n); res["diagonals"] = nd; res["nnz"] = int(p2.count.size); res["host_pixel_build_s"] = round(build_s, 2)
    res["band_bytes"] = int(n) * nd * 8
    res["pixels"]["band_gbs"] = round(res["band_bytes"] / (res["pixels"]["kernel_ms"] * 1e-3) / 1e9, 1)
    res["same"] = len(set(res["check"])) == 1
    out["calculate_expected"] = res
    # chromosome as a trivial assembly
    dp = p2.device_pixels(ctx, "sweight")
    lo, hi = p2.global_bins("chr1", 0, n)
    tt = {}
    for rep in range(2):
        ctx.sync(); ctx.profile_reset()
        t0 = time.perf_counter()
        b = ScoreBatch(4, 400, ctx)
        _job, kept = b.add_pixels(dp, lo, hi, balance=True)
        ctx.sync()
        tt = {"wall_s": round(time.perf_counter() - t0, 4), "kept": int(kept.size),
              "norm_kernel_ms": prof(ctx)[0], "band_kernel_ms": prof(ctx, "bandify")[0]}
        b.close()
    out["chromosome_from_pixels"] = tt
    print(json.dumps(out))


if __name__ == "__main__":
    main()
