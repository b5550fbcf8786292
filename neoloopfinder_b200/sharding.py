"""Sharding of the neoloop-caller work over GPUs / ranks.

The path partitions by assembly with no exchange step: the reference fans
assemblies out to loky processes (scripts/neoloop-caller:155-159) and merges their dictionaries on
the host (:161-170).  Here the units are weighted by their estimated pixel count and dealt to
GPUs with the greedy longest-processing-time rule; results come back as KB-sized dictionaries
(``all_gather_object`` when running under torch.distributed).  No collective on the data path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence


def estimate_pixels(n_bins: int) -> int:
    """Upper bound of scored pixels for a gap-free matrix of n bins."""
    return max(0, 393 * n_bins - 80172) if n_bins >= 401 else max(0, n_bins * (n_bins - 8) // 2)


def assembly_bins(clr, line: str, rsize: int) -> Optional[int]:
    """Exact bin count of the matrix ``complexSV(..., flexible=False)`` builds for an assembly line: geometry only
    (no contact data is read).  None when the line cannot be laid out (the caller falls back to the estimate)."""
    try:
        from .assembly import complexSV
        fu = complexSV(clr, line, span=rsize, col="sweight", expected_values={}, flexible=False)
        fu.reorganize_matrix(map_only=True)
        return int(fu._n_bins)
    except Exception:      # noqa: BLE001 - malformed chains are reported later, by the call itself
        return None


def assembly_weight(line: str, rsize: int, res: int, clr=None) -> float:
    """Cost estimate of one assembly = estimated scored pixels.  With the map at hand the bin count comes from the
    assembly's real geometry (inner blocks of a complex assembly are 0.3-1.5 Mb, far from a fixed guess: on the
    192-assembly bench list the 8-way LPT imbalance of the estimate drops from 3.0 % to 0.03 %); without it a k-SV
    assembly is taken as two outer blocks of rsize and inner ones of rsize/2."""
    if clr is not None:
        n = assembly_bins(clr, line, rsize)
        if n:
            return float(estimate_pixels(n))
    nsv = max(1, len(line.split()) - 2)
    return float(estimate_pixels(int((2 + 0.5 * (nsv - 1)) * rsize // res)))


def lpt_shards(weights: Sequence[float], n_shards: int) -> List[List[int]]:
    """Greedy longest-processing-time partition of item indices into n_shards lists
    (deterministic: ties by index)."""
    order = sorted(range(len(weights)), key=lambda k: (-weights[k], k))
    loads = [0.0] * n_shards
    shards: List[List[int]] = [[] for _ in range(n_shards)]
    for k in order:
        s = min(range(n_shards), key=lambda q: (loads[q], q))
        shards[s].append(k)
        loads[s] += weights[k]
    return shards


def shard_ids(ids: Sequence[str], lines: Dict[str, str], rsize: int, res: int, world: int, clr=None) -> List[List[str]]:
    w = [assembly_weight(lines[i], rsize, res, clr) for i in ids]
    return [[ids[k] for k in shard] for shard in lpt_shards(w, world)]


def gather_dicts(local: Dict, group=None) -> Optional[Dict]:
    """Merge per-rank {assembly ID: result} dictionaries on every rank (host objects, KB-sized)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    out = {}
    for p in parts:
        out.update(p)
    return out


def call_resolution_distributed(clr, assemblies, rsize, balance, prob, mmp, nopool, model_path, expected, stats=None,
                                nproc=1, ctx=None):
    """One resolution under torchrun: every rank (one GPU each) calls its shard, results are
    gathered on all ranks in assembly order."""
    import torch.distributed as dist
    from .cli import call_assemblies
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    ids = list(assemblies)
    mine = shard_ids(ids, assemblies, rsize, clr.binsize, world, clr)[rank]
    local = call_assemblies(clr, mine, assemblies, rsize, balance, prob, mmp, nopool, model_path, expected, ctx=ctx,
                            stats=stats, nproc=nproc)
    merged = gather_dicts(local)
    return [merged[i] for i in ids]
