#!/usr/bin/env python
"""bench.py -- candidate pixels scored per second on the neoloop-caller scoring path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (CPU arm: the oracle port of the reference)

Default workload = the configuration BASELINE.json's metric is quoted on (configs[2], "cfg3"): a synthetic
5 kb hg38-sized contact map with 192 complex assemblies (chains of 2-5 SVs, 3-6 blocks, 1.2-4 k bins each),
3 Mb span.  The list is FIXED: under torchrun the product's own sharder (``sharding.shard_ids``, greedy LPT
on estimated pixels) deals it to the ranks, so 1 -> 8 GPUs is STRONG scaling of one job.  The rank's part
of the sample lives in a pixel table (cooler's layout) resident in HBM; assemblies are built, normalised and
gap-filtered from it by the product path (SURVEY.md section 8 rows a-C .. a-E, f-2).

A "step" is one pass of the scoring path -- gap-free matrix in -> called loops out:
per-diagonal means + isotonic expected, candidate gate, window features, forest, threshold, pooling -- over
the rank's whole shard.  ``value`` times it with the matrices already resident in HBM; ``e2e`` times the
public call ``score_matrices`` with HOST (pinned) matrices, host->device copies and result read-back inside
the timed region.  ``cli_e2e`` is the wall time of the product entry ``call_assemblies`` (what
``neoloop-caller`` runs per resolution) on the same shard: matrix assembly from the pixel table, diagonal
means, every host regression, scoring, pooling and coordinate bookkeeping.  ``parity`` compares, in the
same run, the GPU results with the CPU port of the reference on the sample the CPU baseline is timed on.
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
SPAN = 3_000_000
THRESHOLD = 0.9
SEED = 100
MODEL = os.path.join(ROOT, "tests", "golden", "forest_w6.joblib")
WORKLOADS = {
    # name: (resolution, synth kind, total assemblies, description)
    "cfg3": (5000, "complex", 192, "cfg3 (BASELINE.json configs[2]): synthetic 5 kb hg38-sized map, %d complex assemblies "
                                   "(chains of 2-5 SVs) in total"),
    "cfg2": (10000, "simple", 50, "cfg2 (BASELINE.json configs[1]): synthetic 10 kb hg38-sized map, %d simple SV assemblies in total"),
    "cfg5": (10000, "simple", 50, "cfg5 (BASELINE.json configs[4]): multi-sample batch, SAMPLES synthetic 10 kb hg38-sized maps x %d simple "
                                  "SV assemblies each, (sample, assembly) units dealt to the GPUs"),
    "cfg4": (5000, "genome", 23, "cfg4 (BASELINE.json configs[3]): genome-wide 5 kb call, %d chromosomes of a synthetic hg38-sized map as "
                                 "trivial assemblies"),
}
CROP = 700              # bins of a parity / CPU-baseline matrix (about 15-20 s of one core per matrix at 5 kb)
CPU_SLICE = 50000       # candidates per work unit of the reference arm (about 13 s of one core)


class Workload:
    def __init__(self, name, n_assemblies, n_samples=0):
        self.name = name
        self.res, self.kind, default_n, self.what = WORKLOADS[name]
        self.n = n_assemblies or default_n
        self.n_samples = (n_samples or 50) if name == "cfg5" else 1
        self.what = self.what.replace("SAMPLES", str(self.n_samples))

    def build(self, sample=0):
        from neoloopfinder_b200.synth import analytic_expected, make_workload
        clr, lines = make_workload(self.kind, res=self.res, seed=SEED + sample, n_assemblies=self.n, span=SPAN)
        return clr, lines, analytic_expected(clr)

    def build_samples(self):
        """{sample name: {clr, assemblies, expected}} -- one sample for cfg2 / cfg3, n_samples seeds for cfg5."""
        out = {}
        for k in range(self.n_samples):
            clr, lines, expected = self.build(k)
            out["S%d" % k] = {"clr": clr, "assemblies": lines, "expected": expected}
        return out

    def config(self, n_gpus):
        return {
            "workload": "%s, 3 Mb span, prob %.2f" % (self.what % self.n, THRESHOLD),
            "resolution": self.res, "assemblies_total": self.n * self.n_samples, "samples": self.n_samples, "span_bp": SPAN,
            "forest": "synthetic sklearn RandomForest, 100 trees, depth<=20, 32048 nodes, 169 features (tests/golden/forest_w6.joblib)",
            "expected": "analytic A/(d+1) of the procedural map",
            "source": "regional pixel table (cooler layout) resident in HBM; assemblies built on the device",
            "sharding": ("one fixed list of %s dealt to %d ranks by %s (LPT on estimated pixels), one process per GPU, no "
                         "collective on the data path" % (("(sample, assembly) units", n_gpus, "multisample.shard_sample_units")
                                                          if self.n_samples > 1 else ("assemblies", n_gpus, "sharding.shard_ids")))
                        if n_gpus > 1 else "single GPU",
            "cache": "L2 flushed (256 MB memset) at the start of every step",
        }


def gap_free_matrices(clr, ids, lines, expected, ctx, forest, res, nproc=1):
    """Untimed preparation with the product path (cli.call_assemblies: assemble, normalise, remove gaps):
    -> (gap-free host matrices, block id per gap-free bin) of the valid assemblies, in list order."""
    from neoloopfinder_b200.cli import call_assemblies
    collected = {}
    call_assemblies(clr, ids, lines, SPAN, "sweight", THRESHOLD, 2, False, forest, expected, ctx, nproc=nproc,
                    collect=collected)
    order = [ID for ID in ids if ID in collected]
    return [collected[ID][0] for ID in order], [collected[ID][1] for ID in order]


# --------------------------------------------------------------------------------------------
# CPU side (oracle port): only the cpu_baseline leg and the reference arm come here
# --------------------------------------------------------------------------------------------
def _cpu_score_whole(args):
    """The reference's scoring path (Peakachu init, candidates, getwindow, predict_proba, threshold,
    pooling) on one whole gap-free matrix -> (clist, probas, sorted loop list)."""
    G, chains, expected, model_path, thre, res = args
    from oracle import pyport
    core = pyport.PeakachuPort(G, upper=400 * res, res=res)
    core.load_expected(expected)
    core.load_models(model_path)
    cap = {}
    loops = core.predict(thre=thre, no_pool=False, min_count=2, index_map=None, chains=chains, capture=cap)
    clist = np.asarray(cap.get("clist", np.zeros((0, 2))), dtype=np.int64).reshape(-1, 2)
    probas = np.asarray(cap.get("probas", np.zeros(0)), dtype=np.float64)
    return clist, probas, sorted((int(i), int(j)) for i, j, _p in loops)


_CPU_CORES = {}      # per worker process: matrix -> (PeakachuPort, all candidate rows, all candidate columns)


def _cpu_score_slice(args):
    """Same path restricted to one slice of the matrix's candidate list (reference arm).  The per-matrix
    set-up (CSR, candidates, model load: once per ~200 k candidates in the reference) is kept per worker
    process, so short slices measure the same per-pixel rate as whole matrices."""
    G, chains, expected, model_path, lo, hi, thre, res = args
    from oracle import pyport
    key = (G.shape[0], float(G[::5, ::5].sum()), model_path)
    if key not in _CPU_CORES:
        core = pyport.PeakachuPort(G, upper=400 * res, res=res)
        core.load_expected(expected)
        core.load_models(model_path)
        _CPU_CORES[key] = (core, core.ridx, core.cidx)
    core, ridx, cidx = _CPU_CORES[key]
    core.ridx, core.cidx = ridx[lo:hi], cidx[lo:hi]
    cap = {}
    core.predict(thre=thre, no_pool=False, min_count=2, index_map=None, chains=chains, capture=cap)
    return int(cap.get("n_scored", 0))


def _cpu_prepare(args):
    """Assembly -> gap-free matrix with the oracle port (reference arm)."""
    clr, line, expected = args
    from oracle import pyport
    fu = pyport.AssemblyPort(clr, line, span=SPAN, col="sweight", expected_values=expected, flexible=False)
    fu.reorganize_matrix()
    if len(fu.Map) != fu.fusion_matrix.shape[0]:
        return None
    fu.correct_heterozygosity()
    fu.remove_gaps()
    n = fu.fusion_matrix.shape[0]
    chain_of = np.array([fu.chains[k] for k in range(n)])
    kept = np.array([fu.index_map[k] for k in range(len(fu.index_map))], dtype=np.int64)
    return np.ascontiguousarray(fu.gap_free), chain_of[kept]


def crop_matrix(G, chains, size=CROP):
    """A size x size diagonal sub-matrix around the first junction of the assembly (upper triangular like
    its parent): the unit of the parity check and of the CPU baseline."""
    n = G.shape[0]
    if n <= size:
        return np.ascontiguousarray(G), np.asarray(chains)
    cuts = np.flatnonzero(np.diff(np.asarray(chains)))
    j = int(cuts[0]) + 1 if len(cuts) else n // 2
    a = int(min(max(0, j - size // 2), n - size))
    return np.ascontiguousarray(G[a:a + size, a:a + size]), np.asarray(chains)[a:a + size]


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


def host_cores():
    """Cores this process may really use: loky's count honours CPU affinity and cgroup quotas (os.cpu_count() does not),
    so the host-side worker pools are not oversubscribed on a box that exposes more cores than it grants."""
    try:
        from joblib import cpu_count
        return max(1, int(cpu_count()))
    except Exception:      # noqa: BLE001
        return os.cpu_count() or 1


def n_band_cells(n, upper=400, w=7):
    d = np.arange(0, min(n - 1, upper + 2 * w - 2) + 1)
    return int((n - d).sum())


def dense_upload_bytes(n, bw=414):
    """Bytes nlf_batch_add_dense moves host -> device for an n x n matrix (csrc/score.cu): everything below
    1024 bins, else per block of 128 rows only the columns k_bandify reads."""
    if n < 1024:
        return 8 * n * n
    total = 0
    for r0 in range(0, n, 128):
        r1 = min(n, r0 + 128)
        total += 8 * (min(n, r1 - 1 + bw) - max(0, r0 - 13)) * (r1 - r0)
    return total


# --------------------------------------------------------------------------------------------
def run_reference_arm(args, wl, rank):
    """CPU arm: the oracle port (the reference tree does not travel to the GPU box) on all host
    cores, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    from joblib import Parallel, delayed
    cores = host_cores()
    config = wl.config(args.gpus)
    if wl.name == "cfg4":
        wl = Workload("cfg3", 0)          # the port has no banded input: its pixels/s is measured on 5 kb assembly matrices
    clr, lines, expected = wl.build()
    n_mats = 4 if wl.name == "cfg3" else 6
    todo = [(clr, ln, expected) for ln in list(lines.values())[:n_mats + 2]]
    prepared = Parallel(n_jobs=min(cores, len(todo)))(delayed(_cpu_prepare)(t) for t in todo)
    items = [p for p in prepared if p is not None][:n_mats]
    per_matrix = 4 if wl.name == "cfg3" else 3          # slices of CPU_SLICE candidates a matrix offers
    workers = min(cores, per_matrix * len(items))
    if args.warmup > 0:
        Parallel(n_jobs=workers)(delayed(int)(k) for k in range(workers))       # start the loky workers
    # bounded sample: a step is one work unit per worker; the slice shrinks with --steps so that the whole
    # run stays near three minutes whatever K the caller asks for (about 0.26 ms of one core per candidate)
    cpu_slice = int(min(CPU_SLICE, max(6000, CPU_SLICE * 14 // max(1, args.steps))))
    pix, secs = 0, 0.0
    for _ in range(args.steps):
        units = []
        for k in range(workers):
            G, ch = items[k % len(items)]
            s = k // len(items)
            units.append((G, ch, expected, MODEL, s * cpu_slice, (s + 1) * cpu_slice, THRESHOLD, wl.res))
        t0 = time.perf_counter()
        scored = Parallel(n_jobs=workers)(delayed(_cpu_score_slice)(u) for u in units)
        secs += time.perf_counter() - t0
        pix += sum(scored)
    value = pix / secs
    sample = ("%d work units per step = %d assemblies x candidate slices of %d, oracle/pyport.py, joblib loky n_jobs=%d"
              % (workers, len(items), cpu_slice, workers))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_genome_wide(args, wl, rank, world, local_rank):
    """cfg4: every chromosome of a synthetic hg38-sized 5 kb map as a trivial assembly (genome.call_chromosomes), the
    chromosomes dealt to the ranks by LPT on their bin counts.  A step is one whole call from the pixel table resident
    in HBM (gap removal and bands on the device, K5-K10, coordinates); e2e also uploads the rank's pixel table from
    pinned host memory inside the timed region."""
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; neoloopfinder_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.coolshim import HG38_CHROMSIZES, PixelCooler, SyntheticCooler
    from neoloopfinder_b200.forest import load_forest
    from neoloopfinder_b200.genome import _Subset, call_chromosomes
    from neoloopfinder_b200.scoring import DevicePixels
    from neoloopfinder_b200.sharding import lpt_shards
    from neoloopfinder_b200.synth import analytic_expected

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    ctx = _lib.Context.get(local_rank)
    forest = load_forest(MODEL, ctx)
    names = args.chroms.split(",") if args.chroms else [c for c in HG38_CHROMSIZES if c != "chrY"]
    sizes = {c: HG38_CHROMSIZES[c] for c in names}
    mine = [names[k] for k in sorted(lpt_shards([float(sizes[c]) for c in names], world)[rank])]
    clr = SyntheticCooler(binsize=wl.res, chromsizes=sizes, seed=404, amplitude=255.0 * wl.res / 10000.0)
    expected = analytic_expected(clr)
    t0 = time.perf_counter()
    px = PixelCooler.from_cooler(_Subset(clr, mine), band_bins=449, trans=False) if mine else None
    prep_s = time.perf_counter() - t0

    def one_pass(stats=None):
        return call_chromosomes(px, mine, forest, expected, THRESHOLD, 2, False, "sweight", ctx, stats, 400) if mine else {}

    def sync():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    st = {}
    launches0 = ctx.launch_count()
    calls = one_pass(st)                                        # uploads the table, warms the buffers
    launches_per_pass = ctx.launch_count() - launches0
    for _ in range(max(2, args.warmup - 1)):
        one_pass()
    steps = max(1, args.steps)
    sync()
    ctx.timer_start()
    for _ in range(steps):
        ctx.l2_flush()
        one_pass()
    ms = reduce(max(ctx.timer_stop(), 0.0), dist.ReduceOp.MAX) / steps
    sync()
    # end to end: host pixel table (pinned) -> HBM -> calls, every step
    ms_e, h2d = 0.0, 0
    if mine:
        pin = {k: torch.from_numpy(a).pin_memory() for k, a in (("off", px.bin1_offset), ("b2", px.bin2), ("cnt", px.count),
                                                                 ("w", px._w["sweight"]))}
        h2d = int(sum(t.numel() * t.element_size() for t in pin.values()))
        key = (os.getpid(), ctx.device, "sweight")

        def e2e_pass():
            old = px._device.pop(key, None)
            if old is not None:
                old.close()
            px._device[key] = DevicePixels(pin["off"].numpy(), pin["b2"].numpy(), pin["cnt"].numpy(), pin["w"].numpy(), ctx)
            return one_pass()
        e2e_pass()
    sync()
    ctx.timer_start()
    for _ in range(steps):
        ctx.l2_flush()
        if mine:
            e2e_pass()
    ms_e = reduce(max(ctx.timer_stop(), 0.0), dist.ReduceOp.MAX) / steps
    sync()
    scored = reduce(float(st.get("scored", 0)), dist.ReduceOp.SUM)
    bins = reduce(float(st.get("bins", 0)), dist.ReduceOp.SUM)
    loops = reduce(float(sum(len(v[0]) for v in calls.values())), dist.ReduceOp.SUM)
    h2d_all = reduce(float(h2d), dist.ReduceOp.SUM)
    prep_max = reduce(prep_s, dist.ReduceOp.MAX)
    launches_all = reduce(float(launches_per_pass), dist.ReduceOp.SUM)
    if rank == 0:
        line = {
            "metric": METRIC, "value": scored / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s, prob %.2f" % (wl.what % len(names), THRESHOLD), "resolution": wl.res, "chromosomes": len(names),
                       "gap_free_bins": int(bins), "source": "banded intra-chromosomal pixel table (cooler layout, 449 diagonals) resident "
                       "in HBM; gap removal and bands built on the device",
                       "sharding": "chromosomes dealt to %d ranks by LPT on bin counts, no collective" % world if world > 1 else "single GPU",
                       "cache": "L2 flushed (256 MB memset) at the start of every step"},
            "seconds_per_genome_wide_call": ms * 1e-3,
            "gpu_launches": int(launches_all) * steps,
            "e2e": {"value": scored / (ms_e * 1e-3) if ms_e else None, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all),
                    "d2h_bytes_per_step": int(25 * loops), "ms_per_step": ms_e,
                    "what": "pixel table from pinned host memory -> HBM -> called loops, every step"},
            "pixels": {"scored_per_step_total": int(scored), "loops_per_step_total": int(loops)},
            "prep_seconds_max_over_ranks": {"procedural_map_to_pixel_table": prep_max},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--assemblies", type=int, default=0, help="assemblies per sample (default: 192 for cfg3, 50 for cfg2 / cfg5)")
    ap.add_argument("--samples", type=int, default=0, help="cfg5: number of samples (default 50)")
    ap.add_argument("--chroms", default=None, help="cfg4: comma-separated chromosomes (default: chr1..chr22, chrX)")
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the cli_e2e measurement")
    args = ap.parse_args()
    wl = Workload(args.workload, args.assemblies, args.samples)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return
    if wl.name == "cfg4":
        run_genome_wide(args, wl, rank, world, local_rank)
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
    from neoloopfinder_b200.multisample import call_units, shard_sample_units
    from neoloopfinder_b200.scoring import ScoreBatch, score_and_pool, score_matrices
    from neoloopfinder_b200.sharding import shard_ids
    from neoloopfinder_b200.synth import regional_pixel_cooler

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce(x, dist.ReduceOp.MAX)

    def sum_over_ranks(x):
        return reduce(x, dist.ReduceOp.SUM)

    cores = host_cores()
    nproc = max(1, cores // world)
    ctx = _lib.Context.get(local_rank)
    forest = load_forest(MODEL, ctx)
    t_prep = time.perf_counter()
    syn = wl.build_samples()
    if wl.n_samples == 1:
        # one sample: the product's own sharder deals one fixed assembly list to the ranks
        lines = syn["S0"]["assemblies"]
        units = [("S0", ID) for ID in shard_ids(list(lines), lines, SPAN, wl.res, world, syn["S0"]["clr"])[rank]]
    else:
        units = shard_sample_units(syn, SPAN, world)[rank]          # (sample, assembly) units over the ranks
    mine = {}
    for name, ID in units:
        mine.setdefault(name, []).append(ID)
    expected = syn[next(iter(syn))]["expected"]                     # same analytic table for every seed
    exp_arr = np.r_[[expected[i] for i in sorted(expected)]]
    # this rank's part of every sample it touches, as pixel tables in cooler's layout
    samples = {}
    for name, ids in mine.items():
        s0 = syn[name]
        px = regional_pixel_cooler(s0["clr"], {ID: s0["assemblies"][ID] for ID in ids}, SPAN, nproc=nproc)
        samples[name] = {"clr": px, "assemblies": s0["assemblies"], "model": forest, "expected": s0["expected"]}
    table_s = time.perf_counter() - t_prep

    # ---- the product entry on this shard: warm (pixel table upload, loky workers), then timed wall clock
    cli = None
    my_cli = 0.0
    call_units(samples, units[:2], SPAN, "sweight", THRESHOLD, 2, False, ctx, nproc)
    for s0 in samples.values():
        s0["clr"].device_pixels(ctx, "sweight")
    if not args.no_cli:
        st = {}
        barrier()
        t0 = time.perf_counter()
        res_cli = call_units(samples, units, SPAN, "sweight", THRESHOLD, 2, False, ctx, nproc, st)
        ctx.sync()
        my_cli = time.perf_counter() - t0
        barrier()
        cli_s = max_over_ranks(my_cli)
        cli = {"seconds": cli_s, "assemblies": wl.n * wl.n_samples, "scored_pixels": int(sum_over_ranks(float(st.get("scored", 0)))),
               "loops": int(sum_over_ranks(float(sum(len(r) for r in res_cli.values() if r)))),
               "host_workers_per_gpu": nproc,
               "what": "wall time of the product entry (cli.call_assemblies per sample, multisample.call_units) on every rank's "
                       "shard, max over ranks: assembly from the resident pixel table, diagonal means, host Huber/isotonic "
                       "regressions on loky workers, scoring, pooling, coordinates; the pixel table upload is outside"}
        cli["pixels_per_s"] = cli["scored_pixels"] / cli_s if cli_s > 0 else None
    # ---- untimed preparation with the same product path, keeping the gap-free matrices on the host
    mats, chains = [], []
    for name, ids in mine.items():
        s0 = samples[name]
        m, c = gap_free_matrices(s0["clr"], ids, s0["assemblies"], s0["expected"], ctx, forest, wl.res, nproc)
        mats += m
        chains += c
    pinned = [torch.from_numpy(G).pin_memory() for G in mats]
    host_mats = [t.numpy() for t in pinned]
    del mats
    h2d_bytes = int(sum(dense_upload_bytes(G.shape[0]) for G in host_mats))

    resident = ScoreBatch(4, 400, ctx)
    for G in host_mats:
        resident.add_dense(G)
    ctx.sync()
    prep_s = time.perf_counter() - t_prep

    def step_resident():
        ctx.l2_flush()
        resident.reset()
        return score_and_pool(resident, forest, exp_arr, THRESHOLD, wl.res, 2, chains, False)

    def step_e2e():
        ctx.l2_flush()
        return score_matrices(host_mats, forest, exp_arr, THRESHOLD, wl.res, 2, chains, False, 400, ctx)

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
    ms_mine = max(ctx.timer_stop(), 0.0)
    barrier()
    wall_ms = 1000.0 * (time.perf_counter() - t0)
    launches = sum_over_ranks(float(ctx.launch_count() - launches0))      # whole job, like `value`
    mark1 = sampler.mark()
    ms = max_over_ranks(ms_mine)
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
        step_e2e()
    ms_e_mine = ctx.timer_stop()
    wall_e = 1000.0 * (time.perf_counter() - t_e2e)
    barrier()
    ms_e = max_over_ranks(ms_e_mine)
    clocks = sampler.stop(mark0, mark1)
    e2e_value = total_scored / (ms_e * 1e-3)
    d2h_mine = int(25 * donuts_per_step + 3 * 8 * len(host_mats))
    h2d_total = int(sum_over_ranks(float(h2d_bytes)))
    d2h_total = int(sum_over_ranks(float(d2h_mine)))

    # per-rank table (strong scaling: the slowest rank sets the pace)
    per_rank = None
    if world > 1:
        mine_t = torch.tensor([float(len(host_mats)), float(scored_per_step), ms_mine / args.steps, ms_e_mine / args.steps,
                               my_cli if cli else 0.0], dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(mine_t) for _ in range(world)]
        dist.all_gather(allt, mine_t)
        per_rank = [{"rank": r, "assemblies": int(t[0]), "scored_per_step": int(t[1]), "ms_per_step": float(t[2]),
                     "e2e_ms_per_step": float(t[3]), "cli_seconds": float(t[4])} for r, t in enumerate(allt)]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        nband = sum(n_band_cells(G.shape[0]) for G in host_mats)
        alg_bytes = 8.0 * nband + 16.0 * scored_per_step            # SURVEY.md section 8(d), fused K5-K9, this rank
        kern = {}
        for k, (tms, n) in prof.items():
            if n:
                kern[k] = {"ms_per_step": tms / args.steps, "launches_per_step": n / args.steps}
        feat_ms = prof["feature"][0] / max(1, args.steps)
        forest_ms = prof["forest"][0] / max(1, args.steps)
        score_ms = prof["score"][0] / max(1, args.steps) if "score" in prof else 0.0
        # the dominant kernel by time: forest traversal (LSU / shared-memory bound), features (fp64-issue bound)
        # or, when the two run concurrently on the same SMs, the overlapped pair timed as one region
        if score_ms > 0 and os.environ.get("NLF_SCORE_MODE") == "coop":
            dom, dom_ms = "k_features || k_forest_one (co-resident on every SM, timed together)", score_ms
        elif forest_ms >= feat_ms:
            dom, dom_ms = "k_forest_pipe (all launches of a step)", forest_ms
        else:
            dom, dom_ms = "k_features", feat_ms
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms else 0.0
        fp64_peak = ctx.fp64_peak_gops()
        alg_ops = 4940.0 * scored_per_step                           # add/mul + div per pixel
        band_bytes = 8.0 * sum(G.shape[0] * 416 for G in host_mats)
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if tr.get("workload") == wl.name and dom.startswith("k_forest_pipe"):
                # ncu DRAM bytes of the forest launches per scored pixel (48-assembly capture of the same workload) x pixels
                traffic = tr["k_forest_pipe"]["dram_bytes_per_scored_pixel"] * scored_per_step
        except (OSError, ValueError, KeyError):
            pass
        roofline = {
            "bound": "hbm", "kernel": dom,
            "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes,
            "launch_ms": dom_ms,
            "note": "The contract's HBM fraction on ALGORITHMIC bytes (8*N_band + 16*N_scored, SURVEY.md 8d) for the dominant "
                    "kernel(s).  It is small by construction (SURVEY.md D7: ~29 B but ~5k fp64 ops and ~1600 tree-node visits "
                    "per pixel): the binding resources are the fp64 pipe (roofline_fp64) and the shared-memory wavefront "
                    "rate (roofline_smem); the HBM-bound band kernels are in roofline_band.  traffic = ncu dram bytes of the "
                    "forest launches per scored pixel (profiles/r02_traffic.json, 48-assembly capture) x scored pixels: the "
                    "float32 features (676 B/pixel) are read once per forest part.",
        }
        fp64 = {
            "kernel": "k_features", "achieved_gops": alg_ops / (feat_ms * 1e-3) / 1e9 if feat_ms else None,
            "peak_gops": fp64_peak, "peak_source": "DADD/DMUL issue micro-benchmark in this run (no FMA)",
            "frac": (alg_ops / (feat_ms * 1e-3) / 1e9) / fp64_peak if feat_ms else None,
            "algorithmic_ops_per_pixel": 4940,
            "note": "feature kernel timed over its own launches; when it shares the SMs with the forest kernel its "
                    "launch time includes the slowdown from sharing",
        }
        visits = forest.visits_per_pixel if hasattr(forest, "visits_per_pixel") else 1617.0
        wf = 3.0 * visits * scored_per_step / 32.0
        smem = {
            "kernel": "k_forest", "ms_per_step": forest_ms,
            "node_visits_per_s": visits * scored_per_step / (forest_ms * 1e-3) if forest_ms else None,
            "algorithmic_wavefronts_per_step": wf,
            "achieved_gwavefronts_per_s": wf / (forest_ms * 1e-3) / 1e9 if forest_ms else None,
            "peak_gwavefronts_per_s": 148 * sm_mhz * 1e-3,
            "frac": wf / (forest_ms * 1e-3) / 1e9 / (148 * sm_mhz * 1e-3) if forest_ms else None,
            "peak_source": "148 SMs x sm_max_mhz (MEASURED_PEAKS.json) x 1 shared-memory wavefront per clock",
            "note": "a warp node step needs >= 3 wavefronts (4-byte feature load + 8-byte node load); %.0f node visits "
                    "per pixel on this forest" % visits,
        }
        rows = float(sum(G.shape[0] for G in host_mats))
        band = {}
        # HBM-bound band kernels: bytes they have to read per step / their CUDA-event time
        for k, kname, nbytes in (("diag", "k_diag_means", rows * 397 * 8), ("count", "k_count_candidates", rows * 397 * 8),
                                 ("threshold", "k_threshold", rows * 393 * 8)):
            if k in kern and kern[k]["ms_per_step"] > 0:
                gbs = nbytes / (kern[k]["ms_per_step"] * 1e-3) / 1e9
                band[kname] = {"ms_per_step": kern[k]["ms_per_step"], "algorithmic_bytes_per_step": nbytes, "hbm_gbs": gbs,
                               "frac": gbs / hbm_peak}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": wl.config(world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total,
                    "ms_per_step": ms_e / args.steps, "wall_ms_per_step_rank0": wall_e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_fp64": fp64, "roofline_smem": smem, "roofline_band": band, "kernels": kern,
            "pixels": {"scored_per_step_total": int(total_scored / args.steps), "scored_per_step_rank0": scored_per_step,
                       "candidates_per_step_rank0": cand_per_step, "donuts_per_step_rank0": donuts_per_step,
                       "loops_per_step_rank0": n_loops, "band_cells_rank0": nband, "assemblies_rank0": len(host_mats)},
            "wall_ms_per_step_rank0": wall_ms / args.steps,
            "prep_seconds_rank0": {"pixel_table": table_s, "total_before_timing": prep_s},
            "host": {"os_cpu_count": os.cpu_count(), "usable_cores": cores, "workers_per_rank": nproc},
        }
        if cli:
            line["cli_e2e"] = cli
        if per_rank:
            line["per_rank"] = per_rank
        if not args.no_cpu_baseline and world == 1:                   # reported at N=1 only
            from joblib import Parallel, delayed
            workers = min(cores, 16, len(host_mats))
            picks = [int(round(k * (len(host_mats) - 1) / max(1, workers - 1))) for k in range(workers)]
            crops = [crop_matrix(host_mats[k], chains[k]) for k in picks]
            # GPU on the very same matrices (public batch API), outside every timed region
            pb = ScoreBatch(4, 400, ctx)
            for G, _ch in crops:
                pb.add_dense(G)
            g_loops, _cnt = score_and_pool(pb, forest, exp_arr, THRESHOLD, wl.res, 2, [ch for _G, ch in crops], False)
            g_scored = [pb.scored(j) for j in range(len(crops))]
            pb.close()
            t0 = time.perf_counter()
            cpu = Parallel(n_jobs=workers)(delayed(_cpu_score_whole)((G, ch, expected, MODEL, THRESHOLD, wl.res)) for G, ch in crops)
            secs = time.perf_counter() - t0
            pix = int(sum(len(c[1]) for c in cpu))
            same_pixels, loops_equal, max_dp, flips, n_cmp, n_loops_cmp = True, True, 0.0, 0, 0, 0
            for (gi, gj, gp), gl, (clist, probas, cl) in zip(g_scored, g_loops, cpu):
                if len(gp) != len(probas) or not (np.array_equal(gi, clist[:, 0]) and np.array_equal(gj, clist[:, 1])):
                    same_pixels = False
                    continue
                n_cmp += len(gp)
                if len(gp):
                    max_dp = max(max_dp, float(np.max(np.abs(gp - probas))))
                    flips += int(np.count_nonzero((gp >= THRESHOLD) != (probas >= THRESHOLD)))
                mine_l = sorted(zip(gl[0].tolist(), gl[1].tolist()))
                loops_equal &= (mine_l == cl)
                n_loops_cmp += len(cl)
            line["cpu_baseline"] = {
                "value": pix / secs, "unit": UNIT, "cores": workers, "kind": "port",
                "sample": "%d whole matrices of %d bins (diagonal crops around the first junction of %d assemblies of this "
                          "workload, %d pixels), oracle/pyport.py (reference-faithful Python port: Peakachu init, candidates, "
                          "getwindow, predict_proba, threshold, pooling) over joblib loky n_jobs=%d, %.1f s"
                          % (len(crops), CROP, len(crops), pix, workers, secs)}
            line["parity"] = {"loop_set_equal": bool(loops_equal and same_pixels), "scored_pixel_list_equal": bool(same_pixels),
                              "max_abs_dp": max_dp, "flips": flips, "n_compared": n_cmp, "loops_compared": n_loops_cmp,
                              "threshold": THRESHOLD,
                              "against": "the cpu_baseline run above: same matrices, same forest, GPU through ScoreBatch / score_and_pool"}
        print(json.dumps(line))
    resident.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
