"""Reference-surface behaviours of the drop-in classes beyond the neoloop-caller happy path (SURVEY.md
section 8b): other callers of the same API (``Triangle(Peakachu)``, the ``neotad-caller`` prologue), repeated
``predict`` calls, ``find_anchors`` with any window, worker-process device placement, bundled expected tables."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN, load_case, scale_plan_from_slopes, sha
from tests.test_pyport import rebuild

pytestmark = pytest.mark.gpu


def _core(name="case_transloc_10k"):
    from neoloopfinder_b200.callers import Peakachu
    from oracle import coracle as co
    z = load_case(name)
    op, val = scale_plan_from_slopes(z["slopes"])
    hcm = co.scale_blocks(z["fusion_matrix"], z["block_index"], op, val)
    kept = co.remove_gaps_index(hcm)
    G = hcm[np.ix_(kept, kept)]
    res = int(z["res"])
    core = Peakachu(G, upper=400 * res, res=res)
    core.load_expected({i: v for i, v in enumerate(z["expected"])})
    core.load_models(os.path.join(GOLDEN, str(z["model"])))
    return z, core, G


def test_predict_at_a_second_threshold_reuses_the_scores():
    """callers.py:614-639 may be called again with another threshold: the probabilities stay in HBM and only the
    threshold pass runs again (nlf_batch_rethreshold); the result equals a fresh object's."""
    z, core, G = _core()
    launches = core._ctx.launch_count
    low = sorted(core.predict(thre=0.05, no_pool=True))
    n0 = launches()
    high = sorted(core.predict(thre=float(z["prob"]), no_pool=True))
    assert launches() - n0 <= 2                       # one threshold kernel, no feature / forest launch
    again = sorted(core.predict(thre=0.05, no_pool=True))
    _z, fresh, _G = _core()
    assert high == sorted(fresh.predict(thre=float(z["prob"]), no_pool=True))
    assert again == low and len(low) > len(high) > 0
    assert sorted((i, j) for i, j, _p in high) == sorted(tuple(x) for x in z["donut_ij"])
    pooled = sorted(core.predict(thre=float(z["prob"]), index_map=dict(enumerate(z["index_map"])), chains=dict(enumerate(z["chains"]))))
    assert [(i, j) for i, j, _p in pooled] == [tuple(x) for x in z["loop_index"]]
    clist, probas = core.scored_pixels()
    assert np.array_equal(probas, z["probas"]) and np.array_equal(clist, z["clist"])


@pytest.mark.parametrize("res", [5000, 10000, 25000])
def test_find_anchors_any_window(res):
    """Peakachu.find_anchors(pos, min_count, min_dis, wlen) (callers.py:740-784) for windows other than the default
    40 kb, against the port (scipy find_peaks / peak_widths)."""
    from neoloopfinder_b200.callers import Peakachu
    from oracle import pyport
    rng = np.random.default_rng(res)
    core = object.__new__(Peakachu)
    core.r = res
    port = object.__new__(pyport.PeakachuPort)
    port.r = res
    for trial in range(12):
        centres = rng.integers(50, 400, size=int(rng.integers(2, 9)))
        pos = np.concatenate([np.clip(np.round(rng.normal(c, rng.choice([0.8, 2.0, 5.0]), size=int(rng.integers(3, 40)))), 0, 500)
                              for c in centres]).astype(int)
        for wlen in (20000, 40000, 60000, 100000, 400000):
            if res <= 10000 and min(wlen // res, 20) <= 1:
                with pytest.raises(ValueError):
                    core.find_anchors(pos, min_count=2, min_dis=15000, wlen=wlen)
                continue
            want = port.find_anchors(pos, min_count=2, min_dis=15000, wlen=wlen)
            got = core.find_anchors(pos, min_count=2, min_dis=15000, wlen=wlen)
            assert got == {tuple(int(v) for v in a) for a in want}, (trial, wlen)


def test_triangle_style_subclass_uses_local_clustering():
    """neoloop/visualize/core.py:23, :271, :315: ``class Triangle(Peakachu)`` never runs Peakachu.__init__, sets
    ``self.r`` and calls ``local_clustering(Donuts, min_count=1, r=20000, chains=...)`` without an index map."""
    from neoloopfinder_b200.callers import Peakachu
    from oracle import pyport

    class Triangle(Peakachu):
        def __init__(self, res, chains):
            self.res = res
            self.chains = chains

        def pool(self, Donuts):
            self.r = self.res
            return self.local_clustering(Donuts, min_count=1, r=20000, chains=self.chains)

    z = np.load(os.path.join(GOLDEN, "pool_cases.npz"))
    checked = 0
    for k in range(int(z["n_cases"])):
        res = int(z[f"c{k}_par"][0])
        ij, val, ch = z[f"c{k}_ij"], z[f"c{k}_val"], z[f"c{k}_chain"]
        Donuts = {(int(a), int(b)): float(v) for (a, b), v in zip(ij, val)}
        chains = {int(i): int(c) for i, c in enumerate(ch)}
        port = object.__new__(pyport.PeakachuPort)
        port.r = res
        want = port.local_clustering(dict(Donuts), min_count=1, r=20000, chains=chains)
        got = Triangle(res, chains).pool(Donuts)
        assert got == {(int(a), int(b)) for a, b in want}, k
        checked += 1
        if checked == 20:
            break


@pytest.mark.parametrize("name", ["case_complex_10k", "case_deletion_5k"])
def test_neotad_caller_prologue(name):
    """scripts/neotad-caller:156-170 reuses complexSV(flexible=True) -> reorganize_matrix -> validity check ->
    correct_heterozygosity -> remove_gaps and hands ``gap_free`` (an ndarray) and ``index_to_coordinates`` to its own
    caller.  Same object here; checked against the port on the same map (the flexible assembly is in no fixture)."""
    from neoloopfinder_b200.assembly import complexSV
    from oracle import pyport
    z = load_case(name)
    clr, expected = rebuild(z)
    line = str(z["line"])
    fu = complexSV(clr, line, span=int(z["span"]), col="sweight", protocol="insitu", flexible=True, expected_values=expected)
    fu.reorganize_matrix()
    assert len(fu.Map) == fu.fusion_matrix.shape[0]
    fu.correct_heterozygosity()
    fu.remove_gaps()
    ref = pyport.AssemblyPort(clr, line, span=int(z["span"]), col="sweight", flexible=True, expected_values=expected)
    ref.reorganize_matrix()
    ref.correct_heterozygosity()
    ref.remove_gaps()
    assert fu.gap_free.shape[0] >= 20                                    # neotad-caller:166
    G = np.asarray(fu.gap_free)                                           # what TADcaller(fu.gap_free, ...) would read
    assert isinstance(G, np.ndarray) and sha(G) == sha(ref.gap_free)
    assert {int(k): int(v) for k, v in fu.index_map.items()} == {int(k): int(v) for k, v in ref.index_map.items()}
    tads = [(3, 25, 0), (10, 60, 0)]
    assert fu.index_to_coordinates(tads) == ref.index_to_coordinates(tads)


def test_reference_fanout_with_loky_workers():
    """The reference's own driver loop, ``Parallel(n_jobs=nproc)(delayed(pipeline)(...))`` (scripts/neoloop-caller:
    155-159), over the drop-in pipeline(): every loky worker opens its own context; results equal the in-process call."""
    from joblib import Parallel, delayed
    from neoloopfinder_b200.cli import pipeline
    from neoloopfinder_b200.coolshim import SyntheticCooler
    from neoloopfinder_b200.synth import SMALL_CHROMSIZES, analytic_expected
    svs = [("chr2", 4_000_000, "-", "chr4", 3_000_000, "+"), ("chr1", 6_000_000, "+", "chr1", 8_000_000, "-")]
    clr = SyntheticCooler(binsize=10000, chromsizes=SMALL_CHROMSIZES, seed=21, svs=svs)
    lines = {"C0": "translocation,2,4000000,-,4,3000000,+\t2,4500000\t4,2500000",
             "C1": "deletion,1,6000000,+,1,8000000,-\t1,5500000\t1,8500000"}
    exp = analytic_expected(clr)
    models = {10000: os.path.join(GOLDEN, "forest_w6.joblib")}
    here = [pipeline(clr, k, lines, 600000, "sweight", 0.3, 2, False, models, exp) for k in lines]
    there = Parallel(n_jobs=2)(delayed(pipeline)(clr, k, lines, 600000, "sweight", 0.3, 2, False, models, exp) for k in lines)
    assert there == here and sum(len(r) for r in here) > 0
