#!/usr/bin/env python
"""bench.py -- candidate pixels scored per second on the neoloop-caller scoring path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (CPU arm: the oracle port of the reference)

Workload (BASELINE.json configs[1]): a synthetic 10 kb hg38-sized contact map with 50 simple
translocation / deletion / inversion assemblies per GPU, 3 Mb span.  A "step" is one pass of the
scoring path -- host gap-free matrix in -> called loops out: per-diagonal means +
isotonic expected, candidate gate, window features, forest, threshold, pooling -- over the whole
batch.  ``value`` times it with the matrices already resident in HBM; ``e2e`` times the public
call ``score_matrices`` with HOST (pinned) matrices, host->device copies and result read-back
inside the timed region.  Weak scaling: every rank scores its own 50 assemblies (different seeds).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "candidate pixels scored/sec"
UNIT = "pixels/s"
RES = 10000
SPAN = 3_000_000
N_ASSEMBLIES = 50
THRESHOLD = 0.9
KIND = "simple"             # synth.make_workload kind: 'simple' (cfg2) or 'complex' (cfg3)
WORKLOAD = "cfg2"
MODEL = os.path.join(ROOT, "tests", "golden", "forest_w6.joblib")


def select_workload(name, n_assemblies):
    """cfg2 (default, BASELINE.json configs[1]): 10 kb, simple SV assemblies.  cfg3 (configs[2]): 5 kb, complex
    multi-breakpoint assemblies (2-5 chained SVs, 3-6 blocks, 1.2-4 k bins each); fewer of them per GPU because
    the untimed host preparation fetches every block pair from the procedural map."""
    global RES, KIND, WORKLOAD, N_ASSEMBLIES
    WORKLOAD = name
    if name == "cfg3":
        RES, KIND = 5000, "complex"
    N_ASSEMBLIES = n_assemblies if n_assemblies else (24 if name == "cfg3" else 50)
    return N_ASSEMBLIES


def workload_config(n_gpus):
    what = ("cfg2: synthetic 10 kb hg38-sized map, %d simple SV assemblies per GPU" if WORKLOAD == "cfg2" else
            "cfg3: synthetic 5 kb hg38-sized map, %d complex (2-5 chained SVs) assemblies per GPU") % N_ASSEMBLIES
    return {
        "workload": "%s, 3 Mb span, prob %.2f" % (what, THRESHOLD),
        "resolution": RES, "assemblies_per_gpu": N_ASSEMBLIES, "span_bp": SPAN,
        "forest": "synthetic sklearn RandomForest, 100 trees, depth<=20, 32048 nodes, 169 features (tests/golden/forest_w6.joblib)",
        "expected": "analytic A/(d+1) of the procedural map",
        "sharding": "by assembly, one process per GPU, no collective on the data path" if n_gpus > 1 else "single GPU",
        "cache": "L2 flushed (256 MB memset) at the start of every step",
    }


# --------------------------------------------------------------------------------------------
def build_assemblies(rank, n_asm):
    from neoloopfinder_b200.synth import analytic_expected, make_workload
    clr, lines = make_workload(KIND, res=RES, seed=100 + rank, n_assemblies=n_asm, span=SPAN)
    expected = analytic_expected(clr)
    return clr, lines, expected


def gap_free_matrices_gpu(clr, lines, expected, ctx):
    """Untimed preparation with the product path: assemble, normalise, remove gaps."""
    from neoloopfinder_b200.assembly import complexSV
    mats, chains = [], []
    for ID, line in lines.items():
        fu = complexSV(clr, line, span=SPAN, col="sweight", expected_values=expected, flexible=False)
        fu.reorganize_matrix()
        if len(fu.Map) != fu.fusion_matrix.shape[0]:
            continue
        fu._ctx = ctx
        fu.correct_heterozygosity()
        fu.remove_gaps()
        asm = fu._device_assembly()
        G = asm.gap_free()
        chain_of = np.array([fu.chains[k] for k in range(fu.fusion_matrix.shape[0])])
        chains.append(chain_of[asm.kept_index()])
        mats.append(G)
        asm.close()
    return mats, chains


def gap_free_matrices_cpu(clr, lines, expected, limit):
    """Same preparation with the oracle port (CPU arm)."""
    from oracle import pyport
    out = []
    for ID, line in list(lines.items())[:limit]:
        fu = pyport.AssemblyPort(clr, line, span=SPAN, col="sweight", expected_values=expected, flexible=False)
        fu.reorganize_matrix()
        if len(fu.Map) != fu.fusion_matrix.shape[0]:
            continue
        fu.correct_heterozygosity()
        fu.remove_gaps()
        out.append((fu.gap_free, fu.index_map, fu.chains))
    return out


_CPU_CORES = {}      # per worker process: matrix -> (PeakachuPort, all candidate rows, all candidate columns)


def _cpu_score_one(args):
    """The reference's scoring path (Peakachu init, candidates, getwindow, predict_proba, pooling)
    on one gap-free matrix, restricted to one slice of its candidate list.  The per-matrix set-up
    (CSR, candidates, model load: once per ~150 k candidates in the reference) is kept per worker
    process, so short slices measure the same per-pixel rate as whole matrices."""
    G, index_map, chains, expected, model_path, lo, hi, thre = args
    from oracle import pyport
    key = (G.shape[0], float(G[::5, ::5].sum()), model_path)
    if key not in _CPU_CORES:
        core = pyport.PeakachuPort(G, upper=400 * RES, res=RES)
        core.load_expected(expected)
        core.load_models(model_path)
        _CPU_CORES[key] = (core, core.ridx, core.cidx)
    core, ridx, cidx = _CPU_CORES[key]
    core.ridx, core.cidx = ridx[lo:hi], cidx[lo:hi]
    cap = {}
    core.predict(thre=thre, no_pool=False, min_count=2, index_map=index_map, chains=chains, capture=cap)
    return int(cap.get("n_scored", 0))


CPU_SLICE = 50000      # candidates per work unit (about 13 s of one core)


def cpu_work_units(items, expected, n_units, cpu_slice=CPU_SLICE):
    """(matrix, candidate slice) units: unit k takes slice k // n_items of matrix k % n_items."""
    units = []
    for k in range(n_units):
        G, im, ch = items[k % len(items)]
        s = k // len(items)
        units.append((G, im, ch, expected, MODEL, s * cpu_slice, (s + 1) * cpu_slice, THRESHOLD))
    return units


def cpu_reference_rate(items, expected, workers, n_units=None, cpu_slice=CPU_SLICE):
    """pixels/s of the oracle port over `workers` loky processes (the reference fans out the same
    way, scripts/neoloop-caller:159)."""
    from joblib import Parallel, delayed
    units = cpu_work_units(items, expected, n_units or workers, cpu_slice)
    t0 = time.perf_counter()
    if workers == 1:
        scored = [_cpu_score_one(u) for u in units]
    else:
        scored = Parallel(n_jobs=workers)(delayed(_cpu_score_one)(u) for u in units)
    dt = time.perf_counter() - t0
    return sum(scored) / dt, sum(scored), dt


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        return len(self.lines)

    def stop(self, lo=0, hi=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        time.sleep(0.25)
        hi = len(self.lines) if hi is None else max(hi, lo + 1)
        window = self.lines[max(0, lo - 1):hi + 1] or self.lines
        for line in window:
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def n_band_cells(n, upper=400, w=7):
    d = np.arange(0, min(n - 1, upper + 2 * w - 2) + 1)
    return int((n - d).sum())


# --------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """CPU arm: the oracle port (the reference tree does not travel to the GPU box) on all host
    cores, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    clr, lines, expected = build_assemblies(0, N_ASSEMBLIES)
    items = gap_free_matrices_cpu(clr, lines, expected, 6)
    workers = min(cores, 3 * len(items))           # a 10 kb assembly has ~3 slices of CPU_SLICE candidates
    if args.warmup > 0:
        from joblib import Parallel, delayed
        Parallel(n_jobs=workers)(delayed(int)(k) for k in range(workers))       # start the loky workers
    # bounded sample: a step is one work unit per worker; the slice shrinks with --steps so that the whole
    # run stays near three minutes whatever K the caller asks for (about 0.26 ms of one core per candidate)
    cpu_slice = int(min(CPU_SLICE, max(6000, CPU_SLICE * 14 // max(1, args.steps))))
    pix, secs = 0, 0.0
    for _ in range(args.steps):
        r, p, dt = cpu_reference_rate(items, expected, workers, cpu_slice=cpu_slice)
        pix += p
        secs += dt
    value = pix / secs
    sample = ("%d work units per step = %d assemblies x candidate slices of %d, oracle/pyport.py, joblib loky n_jobs=%d"
              % (workers, len(items), cpu_slice, workers))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--assemblies", type=int, default=0, help="assemblies per GPU (default: 50 for cfg2, 24 for cfg3)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.assemblies = select_workload(args.workload, args.assemblies)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; neoloopfinder_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.forest import load_forest
    from neoloopfinder_b200.scoring import ScoreBatch, score_and_pool, score_matrices

    ctx = _lib.Context.get(local_rank)
    forest = load_forest(MODEL, ctx)
    clr, lines, expected = build_assemblies(rank, args.assemblies)
    exp_arr = np.r_[[expected[i] for i in sorted(expected)]]
    mats, chains = gap_free_matrices_gpu(clr, lines, expected, ctx)
    # pinned host copies for the end-to-end arm
    pinned = [torch.from_numpy(G).pin_memory() for G in mats]
    host_mats = [t.numpy() for t in pinned]
    h2d_bytes = int(sum(G.nbytes for G in host_mats))

    resident = ScoreBatch(4, 400, ctx)
    for G in host_mats:
        resident.add_dense(G)
    ctx.sync()

    def step_resident():
        ctx.l2_flush()
        resident.reset()
        return score_and_pool(resident, forest, exp_arr, THRESHOLD, RES, 2, chains, False)

    def step_e2e():
        ctx.l2_flush()
        return score_matrices(host_mats, forest, exp_arr, THRESHOLD, RES, 2, chains, False, 400, ctx)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- resident (kernel-path) measurement
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(3, args.warmup)):
        loops, counts = step_resident()
    scored_per_step = int(counts[1].sum())
    cand_per_step = int(counts[0].sum())
    donuts_per_step = int(counts[2].sum())
    n_loops = int(sum(len(l[0]) for l in loops))
    ctx.profile(True)
    ctx.profile_reset()
    launches0 = ctx.launch_count()
    barrier()
    mark0 = sampler.mark()
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
    ms = ctx.timer_stop()
    barrier()
    wall_ms = 1000.0 * (time.perf_counter() - t0)
    launches = sum_over_ranks(float(ctx.launch_count() - launches0))      # whole job, like `value`
    mark1 = sampler.mark()
    ms = max_over_ranks(max(ms, 0.0))
    total_scored = sum_over_ranks(float(scored_per_step)) * args.steps
    value = total_scored / (ms * 1e-3)
    prof = {k: ctx.profile_read(v) for k, v in _lib.PROF.items()}
    ctx.profile(False)

    # ---- end-to-end measurement (host matrices, H2D + D2H inside)
    for _ in range(max(3, args.warmup)):
        step_e2e()
    barrier()
    ctx.timer_start()
    t_e2e = time.perf_counter()
    for _ in range(args.steps):
        t_dbg = time.perf_counter()
        loops_e, counts_e = step_e2e()
        if os.environ.get("NLF_BENCH_DEBUG"):
            print("e2e step %.2f ms" % (1e3 * (time.perf_counter() - t_dbg)), file=sys.stderr)
    ms_e = ctx.timer_stop()
    wall_e = 1000.0 * (time.perf_counter() - t_e2e)
    barrier()
    ms_e = max_over_ranks(ms_e)
    clocks = sampler.stop(mark0, mark1)
    e2e_value = total_scored / (ms_e * 1e-3)
    d2h_bytes = int(24 * donuts_per_step + 3 * 8 * len(host_mats) + donuts_per_step)

    if rank == 0:
        # ---- roofline of the dominant kernel
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        nband = sum(n_band_cells(G.shape[0]) for G in host_mats)
        alg_bytes = 8.0 * nband + 16.0 * scored_per_step            # SURVEY.md section 8(d), fused K5-K9
        kern = {}
        for k, (tms, n) in prof.items():
            if n:
                kern[k] = {"ms_per_step": tms / args.steps, "launches_per_step": n / args.steps}
        # dominant kernel = the longest single launch: k_features (one launch per step; the forest runs as
        # 2 part launches + the NaN clean-up).  achieved = algorithmic bytes of the fused scoring path
        #  / that launch's CUDA-event duration.
        feat_ms = prof["feature"][0] / max(1, prof["feature"][1])
        forest_ms = prof["forest"][0] / max(1, args.steps)
        achieved = alg_bytes / (feat_ms * 1e-3) / 1e9
        fp64_peak = ctx.fp64_peak_gops()
        alg_ops = 4940.0 * scored_per_step                           # add/mul + div per pixel
        band_bytes = 8.0 * sum(G.shape[0] * 416 for G in host_mats)
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if WORKLOAD == "cfg2" and tr.get("assemblies") == args.assemblies and tr.get("kernel") == "k_features":
                traffic = tr["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        roofline = {
            "bound": "hbm", "kernel": "k_features",
            "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes,
            "launch_ms": feat_ms,
            "note": "k_features is fp64-issue bound and the forest kernel shared-memory bound by construction "
                    "(SURVEY.md D7: ~29 B but ~5k fp64 ops per pixel), so the HBM fraction on algorithmic bytes is small; "
                    "the fp64-issue fraction is in roofline_fp64. traffic (ncu dram bytes, profiles/r01_traffic.json) exceeds the "
                    "algorithmic bytes because the float32 features (676 B/pixel) go through HBM between k_features and the "
                    "forest kernel.",
        }
        fp64 = {
            "kernel": "k_features", "achieved_gops": alg_ops / (feat_ms * 1e-3) / 1e9 if feat_ms else None,
            "peak_gops": fp64_peak, "peak_source": "DADD/DMUL issue micro-benchmark in this run (no FMA)",
            "frac": (alg_ops / (feat_ms * 1e-3) / 1e9) / fp64_peak if feat_ms else None,
            "algorithmic_ops_per_pixel": 4940,
        }
        forest_info = {
            "kernel": "k_forest_pipe<32 warps, 2 trees per lane> x parts + k_forest_nan", "ms_per_step": forest_ms,
            "node_visits_per_s": 1617.0 * scored_per_step / (forest_ms * 1e-3) if forest_ms else None,
            # shared-memory roofline of the traversal: a warp node step needs at least 3 wavefronts (one
            # conflict-free 4-byte feature load + one 8-byte node load = two 128-byte wavefronts); an SM
            # serves one wavefront per clock
            "smem_roofline": ({
                "algorithmic_wavefronts_per_step": 3.0 * 1617.0 * scored_per_step / 32.0,
                "achieved_gwavefronts_per_s": 3.0 * 1617.0 * scored_per_step / 32.0 / (forest_ms * 1e-3) / 1e9,
                "peak_gwavefronts_per_s": 148 * 1.965,
                "frac": 3.0 * 1617.0 * scored_per_step / 32.0 / (forest_ms * 1e-3) / 1e9 / (148 * 1.965),
                "peak_source": "148 SMs x 1.965 GHz (MEASURED_PEAKS.json sm_max_mhz) x 1 wavefront per clock",
            } if forest_ms else None),
            "note": "about 1617 node visits per pixel (100 trees, mean leaf depth 16.2); bound by shared-memory "
                    "wavefronts and issue slots, see profiles/r01_ncu_summary.md",
        }
        kern_tab = {}
        for k, v in kern.items():
            kern_tab[k] = dict(v)
        if "diag" in kern_tab:
            kern_tab["diag"]["hbm_gbs"] = band_bytes / (kern_tab["diag"]["ms_per_step"] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e / args.steps, "wall_ms_per_step": wall_e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_fp64": fp64, "forest": forest_info, "kernels": kern_tab,
            "pixels": {"scored_per_step_rank0": scored_per_step, "candidates_per_step_rank0": cand_per_step,
                       "donuts_per_step_rank0": donuts_per_step, "loops_per_step_rank0": n_loops,
                       "band_cells_rank0": nband},
            "wall_ms_per_step": wall_ms / args.steps,
        }
        if not args.no_cpu_baseline and world == 1:                   # reported at N=1 only
            cores = os.cpu_count() or 1
            items = [(G, None, ch) for G, ch in zip(host_mats[:8], chains[:8])]
            workers = min(cores, 3 * len(items))
            rate, pix, secs = cpu_reference_rate(items, expected, workers)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": workers, "kind": "port",
                "sample": "%d work units = %d assemblies x candidate slices of %d (%d pixels), oracle/pyport.py "
                          "(reference-faithful Python port) over joblib loky n_jobs=%d, %.1f s"
                          % (workers, len(items), CPU_SLICE, pix, workers, secs)}
        print(json.dumps(line))
    resident.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
