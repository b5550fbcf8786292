"""N > 1 host path on CPU: world_size-2 gloo processes shard the assemblies, "call" their shard
(results looked up from the golden fixtures written by the real reference) and gather; the merged
output must equal the single-process merge.  No GPU involved."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from neoloopfinder_b200.sharding import lpt_shards, shard_ids
from tests.conftest import CASES
from tests.test_pyport import final_from_golden
from tests.helpers import load_case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fixture_results():
    lines, results = {}, {}
    for k, name in enumerate(CASES + ["case_invalid_10k"]):
        z = load_case(name)
        ID = f"C{k}"
        lines[ID] = str(z["line"])
        if bool(z["valid"]):
            results[ID] = {key: (ID, v[1], v[2]) for key, v in final_from_golden(z).items()}
        else:
            results[ID] = None
    return lines, results


def _worker(rank, world, port, queue):
    import torch.distributed as dist
    from neoloopfinder_b200.cli import merge_results
    from neoloopfinder_b200.sharding import gather_dicts
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lines, results = _fixture_results()
    ids = list(lines)
    mine = shard_ids(ids, lines, 3_000_000, 10000, world)[rank]
    local = {i: results[i] for i in mine}
    merged = gather_dicts(local)
    cache = merge_results([merged[i] for i in ids])
    import torch
    n = torch.tensor([float(len(mine))])
    dist.all_reduce(n)
    if rank == 0:
        queue.put((sorted(cache.items()), int(n.item()), mine))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_merge():
    from neoloopfinder_b200.cli import merge_results
    lines, results = _fixture_results()
    ids = list(lines)
    serial = sorted(merge_results([results[i] for i in ids]).items())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, n_total, mine0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == serial
    assert n_total == len(ids)
    assert 0 < len(mine0) < len(ids)


def test_lpt_is_a_balanced_disjoint_cover():
    rng = np.random.default_rng(0)
    w = list(rng.integers(1, 100, size=37).astype(float))
    for world in (1, 2, 4, 8):
        shards = lpt_shards(w, world)
        flat = sorted(k for s in shards for k in s)
        assert flat == list(range(len(w)))
        loads = [sum(w[k] for k in s) for s in shards]
        assert max(loads) - min(loads) <= max(w)


def test_multi_sample_units_are_dealt_once_and_evenly():
    """BASELINE.json configs[4]: (sample, assembly) units of many samples over 8 workers -- every unit exactly once,
    loads within one unit's weight of each other, deterministic."""
    from neoloopfinder_b200.multisample import shard_sample_units
    from neoloopfinder_b200.sharding import assembly_weight
    from neoloopfinder_b200.synth import make_workload

    samples = {}
    for k in range(6):
        clr, lines = make_workload("simple" if k % 2 else "complex", res=10000, seed=50 + k, n_assemblies=7 + k)
        samples["S%d" % k] = {"clr": clr, "assemblies": lines}
    shards = shard_sample_units(samples, 3_000_000, 8)
    flat = [u for s in shards for u in s]
    want = [(n, ID) for n, s in samples.items() for ID in s["assemblies"]]
    assert sorted(flat) == sorted(want) and len(set(flat)) == len(want)
    assert shards == shard_sample_units(samples, 3_000_000, 8)
    w = {(n, ID): assembly_weight(s["assemblies"][ID], 3_000_000, 10000, s["clr"]) for n, s in samples.items() for ID in s["assemblies"]}
    loads = [sum(w[u] for u in s) for s in shards]
    assert max(loads) - min(loads) <= max(w.values())
