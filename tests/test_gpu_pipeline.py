"""End-to-end GPU parity through the reference-shaped API: complexSV / Peakachu / pipeline /
call_assemblies against the golden final dictionaries (real reference) and the Python port run
live on the same procedural map."""
import json
import os

import numpy as np
import pytest

from neoloopfinder_b200.coolshim import SyntheticCooler
from tests.conftest import CASES
from tests.helpers import GOLDEN, load_case
from tests.test_pyport import final_from_golden, rebuild

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_classes_match_golden(name):
    """Drive the drop-in classes exactly like scripts/neoloop-caller:pipeline does, from the
    fixture's stored fusion matrix (independent of the procedural map)."""
    from neoloopfinder_b200.assembly import complexSV
    from neoloopfinder_b200.callers import Peakachu
    z = load_case(name)
    clr, expected = rebuild(z)
    res = int(z["res"])
    fu = complexSV(clr, str(z["line"]), span=int(z["span"]), col=str(z["balance"]), expected_values=expected,
                   flexible=False)
    fu.reorganize_matrix()
    # pin the inputs to the fixture
    fu.fusion_matrix = z["fusion_matrix"].copy()
    fu.bias = z["bias"].copy()
    assert [tuple(x) for x in fu.index] == [tuple(x) for x in z["block_index"]]
    fu._asm = None
    fu.correct_heterozygosity()
    for (a, b), (r, s) in fu.slopes_.items():
        assert r == z["slopes"][a, b, 0] and s == z["slopes"][a, b, 1]
    fu.remove_gaps()
    assert np.array_equal(np.array([fu.index_map[i] for i in range(len(fu.index_map))]), z["index_map"])
    core = Peakachu(fu.gap_free, upper=400 * res, res=res)
    core.load_expected(expected)
    core.load_models(os.path.join(GOLDEN, str(z["model"])))
    assert core.w == int(z["w"])
    assert np.array_equal(core.ridx, z["ridx"]) and np.array_equal(core.cidx, z["cidx"])
    loops = core.predict(thre=float(z["prob"]), no_pool=False, min_count=int(z["min_count"]),
                         index_map=fu.index_map, chains=fu.chains)
    got = sorted((i, j, p) for i, j, p in loops)
    want = [(int(a), int(b), float(p)) for (a, b), p in zip(z["loop_index"], z["loop_prob"])]
    assert got == want
    clist, probas = core.scored_pixels()
    assert np.array_equal(clist, z["clist"]) and np.array_equal(probas, z["probas"])
    coords = fu.index_to_coordinates(loops)
    fu2 = complexSV(clr, str(z["line"]), span=int(z["span"]), col=str(z["balance"]), expected_values=expected)
    fu2.reorganize_matrix(map_only=True)
    final = {k: ("C0", v[0], v[1]) for k, v in coords.items() if (k[0], k[1]) in fu2.Map and (k[3], k[4]) in fu2.Map}
    assert final == final_from_golden(z)
    # host-matrix constructor (no device assembly behind it) gives the same calls
    core2 = Peakachu(np.asarray(fu.gap_free), upper=400 * res, res=res)
    core2.load_expected(expected)
    core2.load_models(os.path.join(GOLDEN, str(z["model"])))
    loops2 = core2.predict(thre=float(z["prob"]), min_count=int(z["min_count"]), index_map=fu.index_map, chains=fu.chains)
    assert sorted(loops2) == got


def test_pipeline_matches_port_live():
    """pipeline() with the reference's signature vs the Python port, same procedural map, a case
    that is in no fixture; includes an invalid assembly (None) and the batched call."""
    from neoloopfinder_b200.cli import call_assemblies, pipeline
    from neoloopfinder_b200.synth import SMALL_CHROMSIZES
    from neoloopfinder_b200.util import calculate_expected
    from oracle import pyport
    svs = [("chr2", 4_000_000, "-", "chr4", 3_000_000, "+"), ("chr1", 6_000_000, "+", "chr1", 8_000_000, "-")]
    clr = SyntheticCooler(binsize=10000, chromsizes=SMALL_CHROMSIZES, seed=21, svs=svs)
    lines = {
        "C0": "translocation,2,4000000,-,4,3000000,+\t2,4500000\t4,2500000",
        "C1": "deletion,1,6000000,+,1,8000000,-\t1,5500000\t1,8500000",
        "C2": "duplication,1,5000000,-,1,5300000,+\t1,5800000\t1,4500000",     # overlapping blocks: invalid
    }
    chroms = list(SMALL_CHROMSIZES)
    exp = calculate_expected(clr, chroms, maxdis=5000000, balance="sweight", nproc=1)
    exp_port = pyport.calculate_expected(clr, chroms, maxdis=5000000, balance="sweight")
    assert all(exp[k] == exp_port[k] for k in exp_port)
    model = {10000: os.path.join(GOLDEN, "forest_w6.joblib")}
    stats = {}
    batch = call_assemblies(clr, list(lines), lines, 600000, "sweight", 0.3, 2, False, model[10000], exp, stats=stats)
    for ID in lines:
        cap = {}
        want = pyport.pipeline(clr, ID, lines, 600000, "sweight", 0.3, 2, False, model, exp_port, capture=cap)
        got = pipeline(clr, ID, lines, 600000, "sweight", 0.3, 2, False, model, exp)
        assert got == want, ID
        assert batch[ID] == want, ID
    assert batch["C2"] is None
    assert stats["scored"] > 1000


def test_no_pool_and_threshold_one():
    from neoloopfinder_b200.callers import Peakachu
    z = load_case("case_transloc_10k")
    from oracle import coracle as co
    from tests.helpers import scale_plan_from_slopes
    op, val = scale_plan_from_slopes(z["slopes"])
    hcm = co.scale_blocks(z["fusion_matrix"], z["block_index"], op, val)
    kept = co.remove_gaps_index(hcm)
    G = hcm[np.ix_(kept, kept)]
    expected = {i: v for i, v in enumerate(z["expected"])}
    core = Peakachu(G, upper=4000000, res=10000)
    core.load_expected(expected)
    core.load_models(os.path.join(GOLDEN, "forest_w6.joblib"))
    raw = core.predict(thre=float(z["prob"]), no_pool=True)
    assert sorted((i, j) for i, j, _ in raw) == sorted(tuple(x) for x in z["donut_ij"])
    assert core.predict(thre=1.5) == []


def test_tiny_matrix_raises_like_reference():
    from neoloopfinder_b200.callers import Peakachu
    with pytest.raises(ValueError):
        Peakachu(np.triu(np.ones((8, 8))), upper=4000000, res=10000)
