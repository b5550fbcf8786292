"""Genome-wide call (BASELINE.json configs[3]): whole chromosomes as trivial assemblies from the
resident pixel table must give exactly what the reference's dense constructor gives on the chromosome's gap-free
matrix (callers.py:527-535) -- checked against the drop-in ``Peakachu`` on the dense matrix and, on the smallest
chromosome, against the CPU port -- and jobs longer than 65 535 rows must score like shorter ones."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu

SIZES = {"chr1": 6_000_000, "chr2": 4_500_000, "chr3": 3_000_000}
MODEL = os.path.join(GOLDEN, "forest_w6.joblib")


@pytest.fixture(scope="module")
def sample():
    from neoloopfinder_b200.coolshim import PixelCooler, SyntheticCooler
    from neoloopfinder_b200.synth import analytic_expected
    clr = SyntheticCooler(binsize=10000, chromsizes=SIZES, seed=77)
    return clr, PixelCooler.from_cooler(clr), analytic_expected(clr)


def test_call_genome_equals_dense_constructor_and_port(sample):
    from neoloopfinder_b200.callers import Peakachu
    from neoloopfinder_b200.genome import call_genome, dense_chromosome, genome_lines
    from oracle import pyport
    clr, px, expected = sample
    stats = {}
    calls = call_genome(px, MODEL, expected, prob=0.5, min_count=2, devices=[0], stats=stats)
    assert list(calls) == list(SIZES) and stats["scored"] > 100000
    total = 0
    for c in SIZES:
        G, keep = dense_chromosome(px, c)
        core = Peakachu(G, upper=400 * 10000, res=10000)
        core.load_expected(expected)
        core.load_models(MODEL)
        want = sorted((int(keep[i]) * 10000, int(keep[j]) * 10000, p) for i, j, p in core.predict(thre=0.5, min_count=2))
        core.close()
        a, b, p = calls[c]
        assert list(zip(a.tolist(), b.tolist(), p.tolist())) == want, c
        total += len(want)
    assert total > 10
    # the CPU port on the smallest chromosome (the reference's Peakachu on the dense gap-free matrix)
    G, keep = dense_chromosome(clr, "chr3")
    port = pyport.PeakachuPort(G, upper=400 * 10000, res=10000)
    port.load_expected(expected)
    port.load_models(MODEL)
    want = sorted((int(keep[i]) * 10000, int(keep[j]) * 10000, float(p)) for i, j, p in port.predict(thre=0.5, min_count=2))
    a, b, p = calls["chr3"]
    assert list(zip(a.tolist(), b.tolist(), p.tolist())) == want
    # a procedural source is turned into a pixel table on the way in
    again = call_genome(clr, MODEL, expected, prob=0.5, min_count=2, chroms=["chr3"], devices=[0])
    assert all(np.array_equal(x, y) for x, y in zip(again["chr3"], calls["chr3"]))
    lines = genome_lines(calls, 10000)
    k = next(iter(lines))
    assert k[0] == k[3] and k[2] - k[1] == 10000 and lines[k][0][1] == k[4] - k[1] and lines[k][0][2] == 0


def test_cli_genome_wide(tmp_path, sample):
    from neoloopfinder_b200.cli import run
    from neoloopfinder_b200.genome import call_genome
    clr, px, expected = sample
    path = os.path.join(tmp_path, "map.npz")
    px.save(path)
    out = os.path.join(tmp_path, "loops.txt")
    run(["-O", out, "-H", path, "--genome-wide", "--chroms", "chr2", "chr3", "--model", MODEL, "--prob", "0.5",
         "--balance-type", "CNV", "--gpus", "0", "--logFile", os.path.join(tmp_path, "log.txt")])
    rows = [ln.split("\t") for ln in open(out).read().splitlines()]
    assert rows and {r[0] for r in rows} <= {"chr2", "chr3"} and all(r[6].endswith(",0") for r in rows)
    # the command line computes its own expected (calculate_expected over the autosomes): same call directly
    from neoloopfinder_b200.util import autosomes, calculate_expected
    exp = calculate_expected(px, autosomes(px.chromnames), maxdis=5000000, balance="sweight")
    calls = call_genome(px, MODEL, exp, prob=0.5, chroms=["chr2", "chr3"], devices=[0])
    assert len(rows) == sum(len(v[0]) for v in calls.values())


def test_jobs_longer_than_65535_rows():
    """A band of 70 000 rows (a chromosome at 2 kb): the rows past 65 535 score exactly like the same rows in a job
    of their own."""
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.coolshim import SyntheticCooler
    from neoloopfinder_b200.forest import load_forest
    from neoloopfinder_b200.scoring import ScoreBatch
    from neoloopfinder_b200.synth import analytic_expected
    ctx = _lib.Context.get(0)
    clr = SyntheticCooler(binsize=10000, chromsizes={"chr1": 30_000_000}, seed=5, gap_frac=0.0)
    small = clr.fetch_band("chr1", 414, "sweight")                       # 3 000 rows
    n_big, lo, hi = 70000, 65200, 69700
    big = small[np.arange(n_big) % small.shape[0]]
    big[-1, 1:] = 0
    for r in range(n_big - 414, n_big):                                   # nothing beyond the matrix edge
        big[r, n_big - r:] = 0
    sub = big[lo:hi].copy()
    for r in range(hi - lo - 414, hi - lo):
        sub[r, hi - lo - r:] = 0
    exp = analytic_expected(clr)
    exp_arr = np.r_[[exp[i] for i in sorted(exp)]]
    forest = load_forest(MODEL, ctx)
    b = ScoreBatch(4, 400, ctx)
    b.add_band(np.ascontiguousarray(big))
    b.add_band(np.ascontiguousarray(sub))
    b.score(forest, exp_arr, 0.5)
    bi, bj, bp = b.scored(0)
    si, sj, sp = b.scored(1)
    b.close()
    assert bi.max() > 69000
    # interior of the sub-job: rows whose windows and columns stay inside it
    sel_b = (bi >= lo + 10) & (bj < hi - 430)
    sel_s = (si >= 10) & (sj < hi - lo - 430)
    got = sorted(zip((bi[sel_b] - lo).tolist(), (bj[sel_b] - lo).tolist(), bp[sel_b].tolist()))
    want = sorted(zip(si[sel_s].tolist(), sj[sel_s].tolist(), sp[sel_s].tolist()))
    assert len(want) > 500000 and got == want
