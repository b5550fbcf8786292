"""Host-side pieces of the drop-in boundary that need no GPU: device placement of worker processes, the bundled
expected tables (neoloop/callers.py:120-134)."""
import os

import numpy as np
import pytest

from oracle.run_reference import REFERENCE_ROOT, reference_available


def test_worker_processes_spread_over_the_gpus(monkeypatch):
    """under the reference's own fan-out (loky / pool worker processes) the device is
    worker number modulo the visible GPUs; NLF_DEVICE and LOCAL_RANK win."""
    import multiprocessing
    from neoloopfinder_b200 import _lib
    monkeypatch.delenv("NLF_DEVICE", raising=False)
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    monkeypatch.setattr(_lib, "device_count", lambda: 4)
    proc = multiprocessing.current_process()
    old = proc.name
    try:
        for name, want in (("LokyProcess-1", 0), ("LokyProcess-2", 1), ("LokyProcess-6", 1), ("SpawnPoolWorker-4", 3), ("MainProcess", 0)):
            proc.name = name
            assert _lib.Context.default_device() == want, name
        monkeypatch.setenv("LOCAL_RANK", "2")
        assert _lib.Context.default_device() == 2
        monkeypatch.setenv("NLF_DEVICE", "3")
        assert _lib.Context.default_device() == 3
    finally:
        proc.name = old



LINE = "translocation,1,5003000,+,2,4000500,-\t1,4200000\t2,4800000"


def _clr(res):
    from neoloopfinder_b200.coolshim import SyntheticCooler
    from neoloopfinder_b200.synth import SMALL_CHROMSIZES
    return SyntheticCooler(binsize=res, chromsizes=SMALL_CHROMSIZES, seed=1)


def test_missing_expected_table_is_an_error_not_an_empty_dict(monkeypatch, tmp_path):
    """With expected_values=None the reference loads its bundled table; without the file this build must not go
    on with {} (every regression would return (0, 0, False) and every assembly would be rejected silently)."""
    import sys
    from neoloopfinder_b200.assembly import complexSV
    monkeypatch.setenv("NLF_EXPECTED_DIR", str(tmp_path))
    monkeypatch.setattr(sys, "path", [p for p in sys.path if os.path.abspath(p or ".") != os.path.abspath(REFERENCE_ROOT)])
    for mod in [m for m in sys.modules if m == "neoloop" or m.startswith("neoloop.")]:
        monkeypatch.delitem(sys.modules, mod)
    with pytest.raises(FileNotFoundError):
        complexSV(_clr(10000), LINE, span=800000, col="sweight")
    # a resolution the reference has no table for gets its empty dictionary, as in the reference
    fu = complexSV(_clr(8000), LINE, span=800000, col="sweight")
    assert fu.expected == {}


@pytest.mark.reference
@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("res", [5000, 10000, 25000])
def test_expected_tables_come_from_the_reference_data_folder(monkeypatch, res):
    import joblib
    from neoloopfinder_b200.assembly import complexSV
    folder = os.path.join(REFERENCE_ROOT, "neoloop", "data")
    monkeypatch.setenv("NLF_EXPECTED_DIR", folder)
    fu = complexSV(_clr(res), LINE, span=800000, col="sweight")
    name = {5000: "5k", 10000: "10k", 25000: "25k"}[res]
    want = joblib.load(os.path.join(folder, "gm.insitu.expected.%s.pkl" % name))
    assert fu.expected == want and len(want) in (1000, 500, 200)


def test_missing_weight_column_and_bad_regions_raise_like_cooler():
    """cooler raises KeyError for a balancing column the bin table lacks and ValueError for inverted or
    out-of-bounds regions; the stand-ins must not fall back to another column or clamp silently."""
    from neoloopfinder_b200.coolshim import ArrayCooler, PixelCooler
    sizes = {"chr1": 1_000_000}
    n = 100
    w = np.full(n, 0.1)
    px = PixelCooler(10000, sizes, np.arange(n + 1, dtype=np.int64), np.arange(n, dtype=np.int32), np.ones(n, np.int32), {"weight": w})
    ar = ArrayCooler(10000, sizes, {("chr1", "chr1"): np.eye(n)}, {"weight": {"chr1": w}})
    for clr in (px, ar):
        assert clr.matrix(balance="weight").fetch(("chr1", 0, 200000)).shape == (20, 20)
        with pytest.raises(KeyError):
            clr.matrix(balance="sweight").fetch(("chr1", 0, 200000))
        with pytest.raises(KeyError):
            clr.bins().fetch(("chr1", 0, 200000))["sweight"]
        with pytest.raises(ValueError):
            clr.matrix(balance=False).fetch(("chr1", 300000, 200000))
        with pytest.raises(ValueError):
            clr.bins().fetch(("chr1", 0, 1_000_001))
        with pytest.raises(ValueError):
            clr.bins().fetch(("chr7", 0, 10))
    with pytest.raises(KeyError):
        px.device_pixels(object(), "sweight")


def test_spawned_worker_with_its_own_pool_returns_at_once():
    """One worker process per GPU (cli.call_resolution, genome.call_genome, multisample.call_samples) may farm its
    regressions out to loky workers.  multiprocessing joins a child's children at exit and idle loky workers only
    leave after 300 s, so such a worker must stop its pool itself (util.shutdown_workers): the executor closes in
    seconds, not minutes."""
    import multiprocessing as mp
    import time
    from concurrent.futures import ProcessPoolExecutor
    from tests.helpers import pool_child
    t0 = time.time()
    with ProcessPoolExecutor(max_workers=2, mp_context=mp.get_context("spawn")) as ex:
        got = [f.result() for f in [ex.submit(pool_child, 6, True) for _ in range(2)]]
    assert got[0] == got[1] > 0
    assert time.time() - t0 < 60
