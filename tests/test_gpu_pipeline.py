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


def test_band_input_equals_dense_input():
    """the banded (n x 414) input form behind the same scoring path."""
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.forest import load_forest
    from neoloopfinder_b200.scoring import ScoreBatch
    from oracle import coracle as co
    from tests.helpers import scale_plan_from_slopes
    z = load_case("case_deletion_5k")
    op, val = scale_plan_from_slopes(z["slopes"])
    hcm = co.scale_blocks(z["fusion_matrix"], z["block_index"], op, val)
    kept = co.remove_gaps_index(hcm)
    G = hcm[np.ix_(kept, kept)]
    n = G.shape[0]
    band = np.zeros((n, 414))
    for d in range(min(414, n)):
        band[:n - d, d] = np.diagonal(G, d)
    ctx = _lib.Context.get(0)
    forest = load_forest(os.path.join(GOLDEN, "forest_w6.joblib"), ctx)
    b = ScoreBatch(4, 400, ctx)
    b.add_dense(G)
    b.add_band(band)
    b.score(forest, z["expected"], float(z["prob"]))
    i0, j0, p0 = b.scored(0)
    i1, j1, p1 = b.scored(1)
    assert np.array_equal(i0, i1) and np.array_equal(j0, j1) and np.array_equal(p0, p1)
    assert np.array_equal(p0, z["probas"])
    b.close()


def test_pipeline_5kb_matches_port_live():
    """5 kb: scipy's peak-distance filter is active (distance 3) and numpy's tie order matters;
    pipeline() must still reproduce the port (which calls scipy/numpy on this machine)."""
    from neoloopfinder_b200.cli import pipeline
    from neoloopfinder_b200.synth import SMALL_CHROMSIZES
    from neoloopfinder_b200.util import calculate_expected
    from oracle import pyport
    sv = ("chr1", 4_000_000, "+", "chr3", 2_500_000, "-")
    clr = SyntheticCooler(binsize=5000, chromsizes=SMALL_CHROMSIZES, seed=31, svs=[sv], amplitude=127.5)
    lines = {"C0": "translocation,1,4000000,+,3,2500000,-\t1,3600000\t3,2900000"}
    exp = calculate_expected(clr, list(SMALL_CHROMSIZES), maxdis=5000000, balance="sweight", nproc=1)
    model = {5000: os.path.join(GOLDEN, "forest_w6.joblib")}
    want = pyport.pipeline(clr, "C0", lines, 450000, "sweight", 0.12, 2, False, model, exp)
    got = pipeline(clr, "C0", lines, 450000, "sweight", 0.12, 2, False, model, exp)
    assert got == want and len(want) > 20


def test_cli_two_resolutions(tmp_path):
    """The neoloop-caller command line end to end (two resolutions, combine_annotations, output
    file) against the Python port driven the same way."""
    from neoloopfinder_b200.cli import run
    from neoloopfinder_b200.synth import SMALL_CHROMSIZES
    from oracle import pyport
    svs = [("chr2", 4_000_000, "-", "chr4", 3_000_000, "+")]
    specs = {}
    for res, amp in ((10000, 255.0), (5000, 127.5)):
        clr = SyntheticCooler(binsize=res, chromsizes=SMALL_CHROMSIZES, seed=41, svs=svs, amplitude=amp)
        path = tmp_path / f"map_{res}.json"
        clr.save(str(path))
        specs[res] = (str(path), clr)
    asm = tmp_path / "test.assemblies.txt"
    asm.write_text("C0\ttranslocation,2,4000000,-,4,3000000,+\t2,4400000\t4,2600000\n")
    out = tmp_path / "loops.txt"
    model = os.path.join(GOLDEN, "forest_w6.joblib")
    run(["-O", str(out), "-H", specs[10000][0], specs[5000][0], "--assembly", str(asm), "-R", "400000",
         "--prob", "0.2", "--model", model, model, "--logFile", str(tmp_path / "log.txt"),
         "--cachefolder", str(tmp_path / "cache")])
    got = [ln.rstrip("\n").split("\t") for ln in open(out)]
    # reference flow with the port
    byres = {}
    lines = {"C0": "translocation,2,4000000,-,4,3000000,+\t2,4400000\t4,2600000"}
    for res in (10000, 5000):
        clr = specs[res][1]
        chroms = [c for c in clr.chromnames]
        exp = pyport.calculate_expected(clr, chroms, maxdis=5000000, balance="sweight")
        r = pyport.pipeline(clr, "C0", lines, 400000, "sweight", 0.2, 2, False, {res: model}, exp)
        cache = {}
        for key in r:
            cache.setdefault(key, []).append(r[key])
        byres[res] = cache
    want = pyport.combine_annotations(byres)
    assert len(want) > 0
    norm = lambda rows: sorted((tuple(r[:6]), tuple(sorted(zip(*[iter(r[6].split(","))] * 3)))) for r in rows)
    assert norm(got) == norm(want)


def test_feature_chunking_gives_identical_results(monkeypatch):
    """A tiny feature-memory budget forces many feature/forest chunks; results must not change."""
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.forest import load_forest
    from neoloopfinder_b200.scoring import ScoreBatch
    from oracle import coracle as co
    from tests.helpers import scale_plan_from_slopes
    z = load_case("case_transloc_10k")
    op, val = scale_plan_from_slopes(z["slopes"])
    hcm = co.scale_blocks(z["fusion_matrix"], z["block_index"], op, val)
    kept = co.remove_gaps_index(hcm)
    G = hcm[np.ix_(kept, kept)]
    ctx = _lib.Context.get(0)
    forest = load_forest(os.path.join(GOLDEN, "forest_w6.joblib"), ctx)
    monkeypatch.setenv("NLF_FEATURE_BUDGET_MB", "1")
    b = ScoreBatch(4, 400, ctx)
    b.add_dense(G)
    b.add_dense(G)
    b.score(forest, z["expected"], float(z["prob"]))
    for job in (0, 1):
        i, j, p = b.scored(job)
        assert np.array_equal(np.stack([i, j], 1), z["clist"])
        assert np.array_equal(p, z["probas"])
    b.close()


def test_long_banded_chromosome_matches_oracle():
    """Config-4 shape: one long chromosome given as its band (n x 414), several feature chunks,
    against the C oracle on the equivalent dense matrix."""
    import joblib
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.forest import load_forest
    from neoloopfinder_b200.scoring import ScoreBatch
    from neoloopfinder_b200.synth import analytic_expected
    from oracle import coracle as co
    n = 2600
    clr = SyntheticCooler(binsize=10000, chromsizes={"chr1": n * 10000}, seed=77)
    G = clr.matrix(balance="sweight").fetch("chr1")
    G[np.isnan(G)] = 0
    G = np.triu(G)
    keep = np.where(G.sum(axis=0) != 0)[0]
    G = np.ascontiguousarray(G[np.ix_(keep, keep)])
    m = G.shape[0]
    band = np.zeros((m, 414))
    for d in range(414):
        band[:m - d, d] = np.diagonal(G, d)
    exp = analytic_expected(clr)
    exp_arr = np.array([exp[i] for i in sorted(exp)])
    ctx = _lib.Context.get(0)
    model_path = os.path.join(GOLDEN, "forest_w6.joblib")
    forest = load_forest(model_path, ctx)
    os.environ["NLF_FEATURE_BUDGET_MB"] = "256"
    try:
        b = ScoreBatch(4, 400, ctx)
        b.add_band(band)
        b.score(forest, exp_arr, 0.9)
        i, j, p = b.scored(0)
        b.close()
    finally:
        del os.environ["NLF_FEATURE_BUDGET_MB"]
    cl, pr, fea32 = co.score_matrix(G, exp_arr, co.sklearn_forest_arrays(joblib.load(model_path)), 6)
    assert len(pr) > 500000
    assert np.array_equal(np.stack([i, j], 1), cl)
    assert np.array_equal(p, pr)
    # the float32 features themselves, bit for bit (8.5e7 cells: a handful of them fall next to a
    # float32 rounding boundary and take the kernel's exact-division path)
    # ... fed as the dense matrix this time: above 1024 bins only the cells near the band are uploaded
    b = ScoreBatch(4, 400, ctx)
    b.keep_features(True)
    b.add_dense(G)
    b.score(forest, exp_arr, 0.9)
    got = b.features32(0)
    i2, j2, p2 = b.scored(0)
    b.close()
    assert got.shape == fea32.shape
    assert np.array_equal(got.view(np.uint32), fea32.view(np.uint32))
    assert np.array_equal(np.stack([i2, j2], 1), cl) and np.array_equal(p2, pr)
    # a non-zero cell below the diagonal is reported, not mis-scored (the trimmed upload still covers
    # the 13 sub-diagonals Peakachu keeps)
    bad = G.copy()
    bad[1500, 1490] = 3.0
    b = ScoreBatch(4, 400, ctx)
    b.add_dense(bad)
    b.score(forest, exp_arr, 0.9)
    with pytest.raises(_lib.NLFError):
        b.scored(0)
    b.close()


def test_calculate_expected_gpu_equals_port():
    """Row f-1: masked per-diagonal sums on the GPU; the expected dictionary must equal the
    reference-order port bit for bit, for a procedural map (band source) and for explicit arrays
    (generic cooler source: sparse scatter path), balanced and raw."""
    from neoloopfinder_b200.coolshim import ArrayCooler
    from neoloopfinder_b200.util import calculate_expected
    from oracle import pyport
    sizes = {"chr1": 6_000_000, "chr2": 4_100_000, "chrX": 3_000_000}
    clr = SyntheticCooler(binsize=10000, chromsizes=sizes, seed=9)
    for balance in ("sweight", "weight", False):
        want = pyport.calculate_expected(clr, ["chr1", "chr2"], maxdis=5000000, balance=balance)
        got = calculate_expected(clr, ["chr1", "chr2"], maxdis=5000000, balance=balance)
        assert list(got) == list(want)
        assert all(got[k] == want[k] for k in want), balance
    # explicit arrays (no fetch_band): small chromosomes, shorter maxdis than the chromosome
    rng = np.random.default_rng(3)
    blocks, weights = {}, {"weight": {}, "sweight": {}}
    small = {"chr1": 1_230_000, "chr2": 800_000}
    for c, size in small.items():
        n = -(-size // 10000)
        B = rng.poisson(3.0, size=(n, n)).astype(float)
        B = np.triu(B) + np.triu(B, 1).T
        blocks[(c, c)] = B
        for col in weights:
            w = rng.uniform(0.05, 0.15, n)
            w[rng.random(n) < 0.05] = np.nan
            weights[col][c] = w
    arr = ArrayCooler(10000, small, blocks, weights)
    want = pyport.calculate_expected(arr, list(small), maxdis=600000, balance="weight")
    got = calculate_expected(arr, list(small), maxdis=600000, balance="weight")
    assert all(got[k] == want[k] for k in want) and len(got) == len(want) == 61


def test_multi_sample_batch_equals_per_sample_calls():
    """BASELINE.json configs[4]: several samples in one run give, per sample, what call_resolution gives alone."""
    from neoloopfinder_b200.cli import call_resolution
    from neoloopfinder_b200.multisample import call_samples
    from neoloopfinder_b200.synth import SMALL_CHROMSIZES, analytic_expected, make_workload
    model = os.path.join(GOLDEN, "forest_w6.joblib")
    samples = {}
    for k in range(3):
        clr, lines = make_workload("simple", res=10000, seed=900 + k, n_assemblies=3, chromsizes=SMALL_CHROMSIZES, span=800000)
        samples["S%d" % k] = {"clr": clr, "assemblies": lines, "model": model, "expected": analytic_expected(clr)}
    stats = {}
    got = call_samples(samples, 800000, "sweight", 0.5, 2, False, devices=[0], stats=stats)
    n = 0
    for name, s in samples.items():
        want = call_resolution(s["clr"], s["assemblies"], 800000, "sweight", 0.5, 2, False, model, s["expected"], devices=[0])
        assert got[name] == want, name
        n += sum(len(r) for r in want if r)
    assert n > 0 and stats["scored"] > 10000


def test_sub_batches_do_not_change_results():
    """call_assemblies works through an assembly file in sub-batches bounded by a device-memory budget (each one is
    closed before the next is built); the split must not change any result."""
    from neoloopfinder_b200.cli import call_assemblies
    from neoloopfinder_b200.synth import SMALL_CHROMSIZES, analytic_expected
    svs = [("chr2", 4_000_000, "-", "chr4", 3_000_000, "+"), ("chr1", 6_000_000, "+", "chr1", 8_000_000, "-")]
    clr = SyntheticCooler(binsize=10000, chromsizes=SMALL_CHROMSIZES, seed=21, svs=svs)
    lines = {
        "C0": "translocation,2,4000000,-,4,3000000,+\t2,4500000\t4,2500000",
        "C1": "deletion,1,6000000,+,1,8000000,-\t1,5500000\t1,8500000",
        "C2": "duplication,1,5000000,-,1,5300000,+\t1,5800000\t1,4500000",     # invalid: costs no device memory
    }
    exp = analytic_expected(clr)
    model = os.path.join(GOLDEN, "forest_w6.joblib")
    s_one, s_many = {}, {}
    one = call_assemblies(clr, list(lines), lines, 600000, "sweight", 0.3, 2, False, model, exp, stats=s_one)
    many = call_assemblies(clr, list(lines), lines, 600000, "sweight", 0.3, 2, False, model, exp, stats=s_many, budget_mb=1)
    assert one == many and one["C2"] is None
    assert s_one["sub_batches"] == 1 and s_many["sub_batches"] == 2
    assert s_one["scored"] == s_many["scored"] > 1000
