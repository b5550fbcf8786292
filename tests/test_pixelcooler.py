"""CPU checks of the pixel-table source: the host view of a
PixelCooler returns the same regions as the source it was made from, survives save/load and
pickling, and -- in the build container -- the REAL reference's complexSV.reorganize_matrix and
_prepare_core give identical results on it."""
import pickle

import numpy as np
import pytest

from neoloopfinder_b200.coolshim import PixelCooler, SyntheticCooler, load_cooler

SIZES = {"chr1": 3_000_000, "chr2": 2_000_000, "chr3": 1_500_000}
SVS = [("chr1", 1_200_000, "+", "chr2", 800_000, "-"), ("chr3", 400_000, "+", "chr3", 900_000, "+")]


@pytest.fixture(scope="module")
def pair():
    clr = SyntheticCooler(binsize=10000, chromsizes=SIZES, seed=3, svs=SVS)
    return clr, PixelCooler.from_cooler(clr)


def test_fetch_equals_source(pair):
    clr, px = pair
    regions = [(("chr1", 0, 1_200_000), ("chr2", 800_000, 2_000_000)),
               (("chr2", 100_000, 900_000), ("chr1", 50_000, 700_000)),          # transposed rectangle
               (("chr1", 0, 3_000_000), ("chr1", 0, 3_000_000)),
               (("chr1", 200_000, 1_000_000), ("chr1", 500_000, 1_500_000)),      # overlapping ranges
               (("chr3", 905_000, 1_500_000), ("chr3", 0, 400_001))]
    for r1, r2 in regions:
        for bal in (False, "weight", "sweight"):
            assert np.array_equal(clr.matrix(balance=bal).fetch(r1, r2), px.matrix(balance=bal).fetch(r1, r2),
                                  equal_nan=True), (r1, r2, bal)
    for c in SIZES:
        for col in ("weight", "sweight"):
            assert np.array_equal(clr.bins().fetch(c)[col].values, px.bins().fetch(c)[col].values, equal_nan=True)
    a = clr.matrix(balance="weight", sparse=True).fetch("chr2").toarray()
    b = px.matrix(balance="weight", sparse=True).fetch("chr2").toarray()
    assert np.array_equal(a, b, equal_nan=True)


def test_layout_and_roundtrip(pair, tmp_path):
    _clr, px = pair
    off, b2 = px.bin1_offset, px.bin2
    assert off[0] == 0 and off[-1] == b2.size and np.all(np.diff(off) >= 0)
    rows = np.repeat(np.arange(off.size - 1), np.diff(off))
    assert np.all(b2 >= rows)                                   # upper triangle
    same_row = rows[1:] == rows[:-1]
    assert np.all(b2[1:][same_row] > b2[:-1][same_row])         # strictly ascending inside a row
    assert px.count.dtype == np.int32 and np.all(px.count > 0)
    path = str(tmp_path / "m.px.npz")
    px.save(path)
    q = load_cooler(path)
    assert isinstance(q, PixelCooler) and q.chromnames == px.chromnames and q.binsize == px.binsize
    assert np.array_equal(q.bin1_offset, off) and np.array_equal(q.bin2, b2) and np.array_equal(q.count, px.count)
    r = pickle.loads(pickle.dumps(px))
    assert np.array_equal(r.matrix(balance="sweight").fetch("chr3"), px.matrix(balance="sweight").fetch("chr3"),
                          equal_nan=True)


def test_band_limited_table():
    clr = SyntheticCooler(binsize=10000, chromsizes={"chr1": 2_000_000, "chr2": 900_000}, seed=5)
    px = PixelCooler.from_cooler(clr, band_bins=50, trans=False)
    full = np.triu(clr.matrix(balance=False).fetch("chr1"))
    got = np.triu(px.matrix(balance=False).fetch("chr1"))
    d = np.subtract.outer(np.arange(full.shape[0]), np.arange(full.shape[0]))
    assert np.array_equal(got, np.where(-d <= 50, full, 0))
    assert not px.matrix(balance=False).fetch("chr1", "chr2").any()


def test_float_counts_are_refused():
    from neoloopfinder_b200.coolshim import ArrayCooler
    B = np.array([[1.5, 0.0], [0.0, 2.0]])
    arr = ArrayCooler(10000, {"chr1": 20000}, {("chr1", "chr1"): B},
                      {"weight": {"chr1": np.ones(2)}, "sweight": {"chr1": np.ones(2)}})
    with pytest.raises(ValueError):
        PixelCooler.from_cooler(arr)


@pytest.mark.reference
def test_real_reference_reads_the_same_from_the_pixel_table(pair):
    from oracle.run_reference import import_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree only exists in the build container")
    _rc, ra, ru, _cli = import_reference()
    clr, px = pair
    lines = ["translocation,1,1200000,+,2,800000,-\t1,600000\t2,1400000",
             "inversion,3,400000,+,3,900000,+\t3,0\t3,500000",
             "translocation,1,1200000,+,2,800000,-\tinversion,2,1500000,+,3,900000,+\t1,600000\t3,300000"]
    for line in lines:
        for col in ("sweight", False):
            a = ra.complexSV(clr, line, span=600000, col=col, protocol="insitu", expected_values={}, flexible=False)
            b = ra.complexSV(px, line, span=600000, col=col, protocol="insitu", expected_values={}, flexible=False)
            a.reorganize_matrix()
            b.reorganize_matrix()
            assert np.array_equal(a.fusion_matrix, b.fusion_matrix) and np.array_equal(a.bias, b.bias)
            assert a.index == b.index and a.Map == b.Map
    for bal in ("sweight", True, False):
        ea = ru._prepare_core((clr, "chr1", 250, bal))
        eb = ru._prepare_core((px, "chr1", 250, bal))
        assert ea.keys() == eb.keys() and all(ea[k][0] == eb[k][0] and ea[k][1] == eb[k][1] for k in ea)


def _blocks_of(px, fu):
    spans = [px.global_bins(*px._bin_extent(tuple(r))) for r in fu.blocks]
    return [a for a, _b in spans], [b for _a, b in spans], [o == '-' for o in fu.orients]


def test_c_oracle_restates_the_pixel_table_reads(pair):
    """oracle/c: reorganize_matrix and _prepare_core on the pixel table equal the Python port reading the
    same table through the cooler surface (and, in the build container, the real reference above)."""
    from neoloopfinder_b200.assembly import complexSV
    from oracle import coracle as co
    from oracle import pyport
    _clr, px = pair
    lines = ["translocation,1,1200000,+,2,800000,-\t1,600000\t2,1400000",
             "inversion,3,400000,+,3,900000,+\t3,0\t3,500000",
             "inversion,3,400000,-,3,900000,-\t3,800000\t3,1300000",
             "translocation,1,1200000,+,2,800000,-\tinversion,2,1500000,+,3,900000,+\t1,600000\t3,300000",
             "duplication,1,1000000,-,1,1200000,+\t1,1500000\t1,700000"]
    for line in lines:
        for col in ("sweight", "weight", False):
            want = pyport.AssemblyPort(px, line, span=600000, col=col, expected_values={}, flexible=False)
            want.reorganize_matrix()
            geo = complexSV(px, line, span=600000, col=col, expected_values={}, flexible=False)
            geo.reorganize_matrix(map_only=True)              # geometry only: no device involved
            assert geo.index == want.index
            lo, hi, rev = _blocks_of(px, geo)
            w = px._w["weight" if col is False else col]
            M, bias = co.assemble_pixels(px.bin1_offset, px.bin2, px.count, w, lo, hi, rev, balance=bool(col))
            assert np.array_equal(M, want.fusion_matrix), (line, col)
            assert np.array_equal(bias, want.bias), (line, col)
    for bal in ("sweight", True, False):
        want = pyport.chrom_expected((px, "chr1", 250, bal))
        lo, hi = px.global_bins("chr1", 0, px._nbins("chr1"))
        sums, cnts = co.expected_chrom(px.bin1_offset, px.bin2, px.count, px._w["weight" if isinstance(bal, bool) else bal],
                                       lo, hi, 251, balance=bool(bal))
        assert set(want) == {i for i in range(251) if cnts[i] > 0}
        assert all(want[i][0] == sums[i] and want[i][1] == cnts[i] for i in want)


def test_c_oracle_pixel_reads_against_dense_numpy():
    """Randomised: small symmetric count matrices with NaN weights; blocks in any order, reversed, overlapping,
    empty; raw and balanced.  The dense construction below is the reference's own recipe (fetch every block
    pair, flip per orientation, paste, np.triu, NaN -> 0; per-diagonal masked sums)."""
    from oracle import coracle as co
    rng = np.random.default_rng(2024)
    for trial in range(60):
        n = int(rng.integers(6, 40))
        Cm = np.triu(rng.poisson(1.2, size=(n, n)))
        Cm = Cm + np.triu(Cm, 1).T                                   # symmetric counts
        w = rng.uniform(0.05, 0.2, n)
        w[rng.random(n) < 0.15] = np.nan
        up = np.triu(Cm)
        rows, cols = np.nonzero(up)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=n), out=off[1:])
        bin2, cnt = cols.astype(np.int32), up[rows, cols].astype(np.int32)
        nblk = int(rng.integers(1, 5))
        lo = rng.integers(0, n, size=nblk)
        hi = np.minimum(n, lo + rng.integers(0, 12, size=nblk))
        rev = rng.random(nblk) < 0.5
        for balance in (True, False):
            S = Cm.astype(float)
            if balance:
                S = S * w[:, None] * w[None, :]
            idx = [np.arange(a, b)[::-1] if r else np.arange(a, b) for a, b, r in zip(lo, hi, rev)]
            order = np.concatenate(idx) if idx else np.zeros(0, int)
            M = S[np.ix_(order, order)]
            # only block pairs j1 <= j2 are pasted, then np.triu
            start = np.r_[0, np.cumsum([len(i) for i in idx])]
            for j1 in range(nblk):
                for j2 in range(j1):
                    M[start[j1]:start[j1 + 1], start[j2]:start[j2 + 1]] = 0
            M = np.triu(M)
            if balance:
                M[np.isnan(M)] = 0
            bias = np.where(np.isnan(w[order]), 0.0, w[order])
            gotM, gotb = co.assemble_pixels(off, bin2, cnt, w, lo, hi, rev, balance=balance)
            assert np.array_equal(gotM, M) and np.array_equal(gotb, bias), (trial, balance)
            # per-diagonal masked sums over the whole "chromosome"
            sums, cnts = co.expected_chrom(off, bin2, cnt, w, 0, n, n, balance=balance)
            valid = np.isfinite(w) & (w > 0)
            for i in range(n):
                v = valid[:n - i] & valid[i:]
                cur = np.diagonal(S, i)[v]
                assert cnts[i] == cur.size and (cur.size == 0 or sums[i] == cur.sum()), (trial, balance, i)
