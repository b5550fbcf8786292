"""Row f-3, host side (no GPU): geometry, matrix, bias and bounds of ``Fusion``, the SV-list parser, the
coordinate maps and the chain formatting / graph helpers of ``assembleSV`` against LIVE calls into the real
reference (build container only; skipped where /root/reference does not exist)."""
import numpy as np
import pytest

from neoloopfinder_b200.coolshim import SyntheticCooler
from oracle.run_reference import reference_available

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")]

SIZES = {"chr1": 12_000_000, "chr2": 9_000_000, "chr3": 7_000_000}
SVS = [("1", 6_000_000, "+", "2", 4_000_000, "-", "translocation"), ("1", 3_000_000, "-", "3", 2_000_000, "-", "translocation"),
       ("2", 2_000_000, "+", "3", 5_000_000, "+", "translocation"), ("3", 1_000_000, "-", "1", 9_000_000, "+", "translocation"),
       ("1", 2_000_000, "+", "1", 5_000_000, "-", "deletion"), ("2", 3_000_000, "+", "2", 3_900_000, "+", "inversion"),
       ("2", 5_000_000, "-", "2", 5_700_000, "-", "inversion"), ("3", 3_000_000, "-", "3", 4_200_000, "+", "duplication")]


@pytest.fixture(scope="module")
def ref():
    from oracle.run_reference import import_reference
    return import_reference()


@pytest.fixture(scope="module")
def clr():
    svs = [("chr" + a, p1, s1, "chr" + b, p2, s2) for a, p1, s1, b, p2, s2, _t in SVS]
    return SyntheticCooler(binsize=10000, chromsizes=SIZES, seed=8, svs=svs)


@pytest.mark.parametrize("sv", SVS)
@pytest.mark.parametrize("trim", [False, True])
def test_fusion_matrix_bias_bounds(ref, clr, sv, trim):
    from neoloopfinder_b200.callers import Fusion
    rc = ref[0]
    c1, p1, s1, c2, p2, s2, note = sv
    for span, col in ((800000, "sweight"), (1500000, "weight")):
        want = rc.Fusion(clr, c1, c2, p1, p2, s1, s2, note, span=span, col=col, trim=trim, expected_values={})
        got = Fusion(clr, c1, c2, p1, p2, s1, s2, note, span=span, col=col, trim=trim, expected_values={})
        assert got.k_p == want.k_p and got.fusion_point == want.fusion_point and got.name == want.name
        assert np.array_equal(got.fusion_matrix, want.fusion_matrix)
        want.extract_bias(col)
        got.extract_bias(col)
        assert np.array_equal(got.bias, want.bias)
        want.detect_bounds()
        got.detect_bounds()
        assert (got.up_i, got.down_i) == (want.up_i, want.down_i)
        assert got.up_var_ratio == want.up_var_ratio and got.down_var_ratio == want.down_var_ratio


def test_stretches_and_locate(ref):
    from neoloopfinder_b200.callers import Fusion
    rc = ref[0]
    a, b = object.__new__(Fusion), object.__new__(rc.Fusion)
    rng = np.random.default_rng(1)
    for _ in range(300):
        n = int(rng.integers(12, 80))
        v = np.cumsum(rng.standard_normal(n)) * (rng.random(n) > 0.1)
        for left in (True, False):
            try:
                want = b.locate(v.copy(), left_most=left)
            except IndexError:
                with pytest.raises(IndexError):
                    a.locate(v.copy(), left_most=left)
                continue
            assert a.locate(v.copy(), left_most=left) == want


def test_sv_list_and_coordinate_maps(ref, tmp_path):
    from neoloopfinder_b200.util import load_translocation_fil, map_coordinates
    ru = ref[2]
    lines = ["chr1 chr2 +- 6000000 4000000 translocation", "chr2 chr2 ++ 3900000 3000000 inversion",
             "chr3 chr3 -+ 3000000 3400000 duplication", "chr3 chr3 -+ 3000000 4200000 duplication",
             "chr1 chr1 +- 2000000-2004000 5000000 deletion", "chr1 chr1 +- 2000000-2100000 5000000 deletion",
             "chr1 chr1 +- 2000000 2300000 deletion", "chr1 chr2 +- 6000000 4000000 translocation",
             "1 3 -- 3000000 2000000-2000010 translocation"]
    path = str(tmp_path / "svs.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    assert load_translocation_fil(path, 10000, 500000) == ru.load_translocation_fil(path, 10000, 500000)
    assert load_translocation_fil(path, 5000, 100000) == ru.load_translocation_fil(path, 5000, 100000)
    for k_p in ([1000000, 2005000, 3000000, 3400000], [2000000, 1000000, 3400000, 3000000],
                [0, 95000, 995000, 400000], [600000, 0, 700000, 1200000]):
        for unit in ("bp", "bin"):
            assert map_coordinates(k_p, 10000, "1", "2", unit) == ru.map_coordinates(k_p, 10000, "1", "2", unit)
            assert map_coordinates(k_p, 10000, "1", "1", unit) == ru.map_coordinates(k_p, 10000, "1", "1", unit)


def test_graph_helpers_match(ref, clr, monkeypatch):
    """Graph construction, chain search and formatting with the contact-map tests stubbed by the same
    deterministic rule on both sides."""
    import neoloopfinder_b200.assembly as mine
    ra = ref[1]
    queue = [
        [("1", 6000000, "+", "2", 4000000, "-", "translocation"), ("1", 5400000, 6010000), ("2", 4000000, 4700000)],
        [("2", 4500000, "+", "3", 2000000, "-", "translocation"), ("2", 3900000, 4510000), ("3", 2000000, 2600000)],
        [("3", 2400000, "+", "1", 9000000, "-", "translocation"), ("3", 1800000, 2410000), ("1", 9000000, 9500000)],
        [("1", 9300000, "+", "2", 1000000, "+", "translocation"), ("1", 8800000, 9310000), ("2", 1010000, 500000)],
        [("3", 5000000, "-", "2", 8000000, "-", "translocation"), ("3", 5600000, 5000000), ("2", 8000000, 8600000)],
    ]
    rule = lambda line: (sum(map(ord, line)) % 11) != 0          # noqa: E731

    def stub(clr_, ID, span, col, path, protocol, expected):
        return rule(ID), path
    monkeypatch.setattr(mine, "filterAssembly", stub)
    monkeypatch.setattr(ra, "filterAssembly", stub)
    outs = []
    for mod in (mine, ra):
        w = object.__new__(mod.assembleSV)
        w.clr, w.res, w.span, w.balance_type, w.protocol, w.expected, w.n_jobs = clr, 10000, 800000, "sweight", "insitu", {}, 1
        w.queue = [list(q) for q in queue]
        w.lookup = {q[0]: [q[1], q[2]] for q in queue}
        w.build_graph()
        w.find_complexSV()
        outs.append((sorted(w.graph.edges(data="weight")), w.alleles,
                     [w._change_format(p) for p in w.alleles], w._containloop(w.alleles[0]) if w.alleles else None))
    assert outs[0] == outs[1]
    assert len(outs[0][0]) >= 4 and len(outs[0][1]) >= 1 and max(len(p) for p in outs[0][1]) >= 3
