"""Many samples in one run (BASELINE.json configs[4]: 50 cancer Hi-C maps x their SV assemblies on 8 GPUs).

The reference handles one sample per invocation (scripts/neoloop-caller:144-170) and fans its assemblies out to
loky processes.  Here the unit of work is one (sample, assembly) pair: all units of all samples are weighted by
their estimated pixel count and dealt to the GPUs with the same longest-processing-time rule as a single sample
(``sharding.lpt_shards``), so a batch of small samples fills every GPU evenly; a GPU walks its units sample by
sample (one model / pixel-table upload per sample it touches) through ``cli.call_assemblies``.  Results are
per-assembly dictionaries gathered on the host; there is no collective.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

from .sharding import assembly_weight, lpt_shards

__all__ = ["shard_sample_units", "call_units", "call_samples"]

Unit = Tuple[str, str]          # (sample name, assembly ID)


def shard_sample_units(samples: Dict[str, dict], rsize: int, n_shards: int) -> List[List[Unit]]:
    """Deal every (sample, assembly) unit to ``n_shards`` workers (LPT on estimated pixels, deterministic)."""
    units, weights = [], []
    for name, s in samples.items():
        res = s["clr"].binsize
        for ID, line in s["assemblies"].items():
            units.append((name, ID))
            weights.append(assembly_weight(line, rsize, res, s["clr"]))
    shards = lpt_shards(weights, n_shards)
    return [[units[k] for k in sorted(shard)] for shard in shards]


def call_units(samples: Dict[str, dict], units: Sequence[Unit], rsize, balance, prob, mmp, nopool, ctx=None, nproc=1,
               stats=None) -> Dict[Unit, Optional[dict]]:
    """The listed units on ONE GPU, grouped by sample.  ``samples[name]`` = {clr, assemblies, model, expected}."""
    from .cli import call_assemblies
    by_sample: Dict[str, List[str]] = {}
    for name, ID in units:
        by_sample.setdefault(name, []).append(ID)
    out = {}
    for name, ids in by_sample.items():
        s = samples[name]
        r = call_assemblies(s["clr"], ids, s["assemblies"], rsize, balance, prob, mmp, nopool, s["model"], s["expected"],
                            ctx, stats, nproc)
        for ID in ids:
            out[(name, ID)] = r[ID]
    return out


def _sample_worker(dev, samples, units, rsize, balance, prob, mmp, nopool, nproc, child=False):
    from . import _lib
    try:
        st = {}
        r = call_units(samples, units, rsize, balance, prob, mmp, nopool, _lib.Context.get(dev), nproc, st)
        return r, st
    finally:
        if child:                      # a spawned worker stops its own loky pool before returning (util.shutdown_workers)
            from .util import shutdown_workers
            shutdown_workers()


def call_samples(samples: Dict[str, dict], rsize, balance, prob, mmp, nopool, devices: Optional[List[int]] = None,
                 nproc=1, stats=None, workers: Optional[str] = None) -> Dict[str, List[Optional[dict]]]:
    """Every assembly of every sample over the visible GPUs.  Returns {sample: [pipeline() result per assembly, in
    the order of its assembly file]} -- what ``call_resolution`` returns for each sample on its own."""
    from . import _lib
    if devices is None:
        devices = list(range(_lib.device_count()))
    devices = list(dict.fromkeys(int(d) for d in devices))
    if not devices:
        raise RuntimeError("no CUDA device visible: neoloopfinder_b200 has no CPU fallback")
    shards = shard_sample_units(samples, rsize, len(devices))
    jobs = [(d, u) for d, u in zip(devices, shards) if u]
    per_dev = max(1, nproc // max(1, len(jobs)))
    mode = workers or os.environ.get("NLF_WORKERS", "process")
    if len(jobs) <= 1:
        parts = [_sample_worker(d, samples, u, rsize, balance, prob, mmp, nopool, per_dev) for d, u in jobs]
    elif mode == "thread":
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            parts = list(ex.map(lambda j: _sample_worker(j[0], samples, j[1], rsize, balance, prob, mmp, nopool, per_dev), jobs))
    else:
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(max_workers=len(jobs), mp_context=mp.get_context("spawn")) as ex:
            futs = [ex.submit(_sample_worker, d, {n: samples[n] for n in {name for name, _ in u}}, u, rsize, balance, prob,
                              mmp, nopool, per_dev, True) for d, u in jobs]
            parts = [f.result() for f in futs]
    merged = {}
    for r, st in parts:
        merged.update(r)
        if stats is not None:
            for k, v in st.items():
                stats[k] = stats.get(k, 0) + v
    return {name: [merged[(name, ID)] for ID in s["assemblies"]] for name, s in samples.items()}
