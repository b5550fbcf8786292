"""SURVEY.md section 8, row f-2 on the GPU: the sample's pixel table resident in HBM.  The matrix
of every assembly (complexSV.reorganize_matrix), the prologue sums (util._prepare_core) and the
gap-free band of whole chromosomes are built on the device and must equal, bit for bit, what the
Python port (oracle/pyport.py: the reference's call sites) reads through the cooler surface."""
import os

import numpy as np
import pytest

from neoloopfinder_b200.coolshim import PixelCooler, SyntheticCooler
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu

SIZES = {"chr1": 12_000_000, "chr2": 9_000_000, "chr3": 7_000_000, "chr4": 6_000_000}
SVS = [("chr2", 4_000_000, "-", "chr4", 3_000_000, "+"), ("chr1", 6_000_000, "+", "chr1", 8_000_000, "-"),
       ("chr3", 2_000_000, "+", "chr3", 5_000_000, "+"), ("chr4", 2_000_000, "-", "chr1", 2_000_000, "+")]
LINES = {
    "C0": "translocation,2,4000000,-,4,3000000,+\t2,4500000\t4,2500000",
    "C1": "deletion,1,6000000,+,1,8000000,-\t1,5500000\t1,8500000",
    "C2": "inversion,3,2000000,+,3,5000000,+\t3,1500000\t3,4500000",
    "C3": "inversion,3,2000000,-,3,5000000,-\t3,2500000\t3,5500000",
    "A0": "translocation,2,4000000,-,4,3000000,+\ttranslocation,4,2000000,-,1,2000000,+\t2,4500000\t1,1500000",
    "C4": "duplication,1,5000000,-,1,5300000,+\t1,5800000\t1,4500000",      # overlapping blocks (invalid assembly)
}


@pytest.fixture(scope="module")
def maps():
    clr = SyntheticCooler(binsize=10000, chromsizes=SIZES, seed=21, svs=SVS)
    return clr, PixelCooler.from_cooler(clr)


@pytest.mark.parametrize("ID", list(LINES))
@pytest.mark.parametrize("col", ["sweight", "weight", False])
def test_assembly_matrix_from_pixels(maps, ID, col):
    from neoloopfinder_b200.assembly import complexSV
    from oracle import pyport
    _clr, px = maps
    want = pyport.AssemblyPort(px, LINES[ID], span=600000, col=col, expected_values={}, flexible=False)
    want.reorganize_matrix()
    fu = complexSV(px, LINES[ID], span=600000, col=col, expected_values={}, flexible=False)
    fu.reorganize_matrix()
    assert fu._asm is not None and fu._fusion_matrix is None          # built on the device, nothing fetched yet
    assert fu.index == want.index and fu.Map == want.Map
    assert np.array_equal(fu.bias, want.bias)
    assert np.array_equal(fu.fusion_matrix, want.fusion_matrix)
    assert fu.fusion_matrix.any()
    # and the plain-C restatement on the same pixel table
    from oracle import coracle as co
    spans = [px.global_bins(*px._bin_extent(tuple(r))) for r in fu.blocks]
    M, bias = co.assemble_pixels(px.bin1_offset, px.bin2, px.count, px._w["weight" if col is False else col],
                                 [a for a, _b in spans], [b for _a, b in spans], [o == '-' for o in fu.orients],
                                 balance=bool(col))
    assert np.array_equal(fu.fusion_matrix, M) and np.array_equal(fu.bias, bias)


def test_pipeline_from_pixels_equals_port(maps):
    from neoloopfinder_b200.cli import call_assemblies
    from neoloopfinder_b200.util import calculate_expected
    from oracle import pyport
    clr, px = maps
    chroms = list(SIZES)
    exp = calculate_expected(px, chroms, maxdis=5000000, balance="sweight", nproc=1)
    exp_port = pyport.calculate_expected(clr, chroms, maxdis=5000000, balance="sweight")
    assert list(exp) == list(exp_port) and all(exp[k] == exp_port[k] for k in exp_port)
    model = {10000: os.path.join(GOLDEN, "forest_w6.joblib")}
    got = call_assemblies(px, list(LINES), LINES, 600000, "sweight", 0.3, 2, False, model[10000], exp)
    n_loops = 0
    for ID in LINES:
        want = pyport.pipeline(clr, ID, LINES, 600000, "sweight", 0.3, 2, False, model, exp_port)
        assert got[ID] == want, ID
        n_loops += len(want or {})
    assert got["C4"] is None and n_loops > 20


@pytest.mark.parametrize("balance", ["sweight", True, False])
def test_expected_from_pixels(balance):
    """Chromosomes of 40, 300 and 6000 bins: diagonals shorter than one numpy leaf, a few leaves, and
    more leaves than the 32 subtrees the block splits into; maxdis longer and shorter than the chromosome."""
    from neoloopfinder_b200.util import calculate_expected
    from oracle import pyport
    sizes = {"chr1": 60_000_000, "chr2": 3_000_000, "chr3": 400_000, "chrX": 1_000_000}
    clr = SyntheticCooler(binsize=10000, chromsizes=sizes, seed=4, band_bp=2_100_000)
    px = PixelCooler.from_cooler(clr, band_bins=205, trans=False)
    chroms = ["chr1", "chr2", "chr3"]
    for maxdis in (2_000_000, 350_000):
        want = pyport.calculate_expected(px, chroms, maxdis=maxdis, balance=balance)
        got = calculate_expected(px, chroms, maxdis=maxdis, balance=balance)
        assert list(got) == list(want) and len(want) == maxdis // 10000 + 1
        assert all(got[k] == want[k] for k in want), (balance, maxdis)
    # raw per-chromosome sums and counts against the plain-C restatement (6 000-bin and 300-bin chromosomes)
    from oracle import coracle as co
    col = "weight" if isinstance(balance, bool) else balance
    dp = px.device_pixels(None, col)
    for c in ("chr1", "chr2"):
        lo, hi = px.global_bins(c, 0, px._nbins(c))
        sums, cnts = dp.expected_diagonals(lo, hi, 201, balance=bool(balance))
        ws, wc = co.expected_chrom(px.bin1_offset, px.bin2, px.count, px._w[col], lo, hi, 201, balance=bool(balance))
        assert np.array_equal(cnts, wc) and np.array_equal(sums, ws), (balance, c)


def test_chromosomes_from_pixels_equal_host_band():
    """Genome-wide shape (config 4): chromosomes as trivial assemblies straight from the pixel table
    vs the host construction (np.triu, remove_gaps, band) fed through add_band."""
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.forest import load_forest
    from neoloopfinder_b200.scoring import ScoreBatch
    from neoloopfinder_b200.synth import analytic_expected
    sizes = {"chr1": 21_000_000, "chr2": 6_000_000, "chr3": 80_000}
    clr = SyntheticCooler(binsize=10000, chromsizes=sizes, seed=77, gap_frac=0.05)
    px = PixelCooler.from_cooler(clr, trans=False)
    exp = analytic_expected(clr)
    exp_arr = np.array([exp[i] for i in sorted(exp)])
    ctx = _lib.Context.get(0)
    forest = load_forest(os.path.join(GOLDEN, "forest_w6.joblib"), ctx)
    dp = px.device_pixels(ctx, "sweight")
    a, b = ScoreBatch(4, 400, ctx), ScoreBatch(4, 400, ctx)
    kept_dev = []
    for c in ("chr1", "chr2"):
        G = clr.matrix(balance="sweight").fetch(c)
        G[np.isnan(G)] = 0
        G = np.triu(G)
        keep = np.where(G.sum(axis=0) != 0)[0]
        if 11 > keep.size:
            keep = np.arange(G.shape[0])
        G = G[np.ix_(keep, keep)]
        m = G.shape[0]
        band = np.zeros((m, 414))
        for d in range(min(414, m)):
            band[:m - d, d] = np.diagonal(G, d)
        a.add_band(band)
        lo, hi = px.global_bins(c, 0, px._nbins(c))
        _job, kept = b.add_pixels(dp, lo, hi, balance=True)
        assert np.array_equal(kept, keep), c
        kept_dev.append(kept)
    a.score(forest, exp_arr, 0.5)
    b.score(forest, exp_arr, 0.5)
    total = 0
    for job in range(2):
        ia, ja, pa = a.scored(job)
        ib, jb, pb = b.scored(job)
        assert np.array_equal(ia, ib) and np.array_equal(ja, jb) and np.array_equal(pa, pb)
        da, db = a.donuts(job), b.donuts(job)                       # appended in any order
        oa, ob = np.lexsort((da[1], da[0])), np.lexsort((db[1], db[0]))
        assert all(np.array_equal(x[oa], y[ob]) for x, y in zip(da, db))
        total += len(pa)
    assert total > 500000
    a.close()
    b.close()
    # fewer than 11 surviving bins: every bin is kept (callers.py:514-516) and the isotonic fit of
    # get_candidates has no input, as in the reference
    t = ScoreBatch(4, 400, ctx)
    lo, hi = px.global_bins("chr3", 0, px._nbins("chr3"))
    _job, kept = t.add_pixels(dp, lo, hi, balance=True)
    assert np.array_equal(kept, np.arange(8))
    t.score(forest, exp_arr, 0.5)
    with pytest.raises(_lib.NLFError):
        t.counts()
    t.close()


def test_pixel_table_is_validated():
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.scoring import DevicePixels
    off = np.array([0, 2, 3, 3], dtype=np.int64)
    w = np.ones(3)
    DevicePixels(off, [0, 2, 1], [5, 1, 2], w).close()
    with pytest.raises(_lib.NLFError):
        DevicePixels(off, [2, 0, 1], [5, 1, 2], w)                 # row 0 not ascending
    with pytest.raises(_lib.NLFError):
        DevicePixels(off, [0, 2, 0], [5, 1, 2], w)                 # below the diagonal
    with pytest.raises(_lib.NLFError):
        DevicePixels(off, [0, 3, 1], [5, 1, 2], w)                 # beyond the last bin
    with pytest.raises(ValueError):
        DevicePixels(off, [0, 2], [5, 1], w)


def test_random_block_layouts_match_c_oracle():
    """k_px_assemble on arbitrary block tables (any order on the genome, reversed, overlapping, empty) against the
    plain-C restatement; raw and balanced, NaN weights."""
    from neoloopfinder_b200.scoring import DeviceAssembly, DevicePixels
    from oracle import coracle as co
    rng = np.random.default_rng(77)
    for trial in range(40):
        n = int(rng.integers(8, 300))
        Cm = np.triu(rng.poisson(0.8, size=(n, n)))
        w = rng.uniform(0.05, 0.2, n)
        w[rng.random(n) < 0.1] = np.nan
        rows, cols = np.nonzero(Cm)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=n), out=off[1:])
        bin2, cnt = cols.astype(np.int32), Cm[rows, cols].astype(np.int32)
        nblk = int(rng.integers(1, 6))
        lo = rng.integers(0, n, size=nblk)
        hi = np.minimum(n, lo + rng.integers(0, 80, size=nblk))
        if int((hi - lo).sum()) == 0:
            hi[0] = min(n, lo[0] + 3)
        rev = rng.random(nblk) < 0.5
        px = DevicePixels(off, bin2, cnt, w)
        for balance in (True, False):
            asm = DeviceAssembly.from_pixels(px, lo, hi, rev, balance=balance)
            M, bias = co.assemble_pixels(off, bin2, cnt, w, lo, hi, rev, balance=balance)
            assert np.array_equal(asm.fusion_matrix(), M), (trial, balance)
            assert np.array_equal(asm.bias(), bias), (trial, balance)
            asm.close()
        px.close()
