"""Row f-3 on the GPU: the inner loop of assemble-complexSVs (Fusion per SV, the SV graph, filterAssembly per
edge and per chain) against the fixture written by the REAL reference (oracle/gen_golden_assemble.py):
per-SV bounds, slopes and R^2 bit for bit, the same surviving SVs, graph edges, alleles and assembly file --
from the procedural map through the cooler surface and from the pixel table resident in HBM."""
import hashlib
import json
import os

import numpy as np
import pytest

from neoloopfinder_b200.coolshim import PixelCooler, SyntheticCooler
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu

SCALARS = ("u_g_r", "u_g_s", "d_g_r", "d_g_s", "m_g_r", "m_g_s", "u_s", "d_s")


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    z = json.load(open(os.path.join(GOLDEN, "assemble_10k.json")))
    clr = SyntheticCooler.from_spec(z["cooler_spec"])
    d = tmp_path_factory.mktemp("asm")
    sv_file = str(d / "svs.txt")
    open(sv_file, "w").write("\n".join(z["sv_lines"]) + "\n")
    spec_file = str(d / "map.json")
    clr.save(spec_file)
    expected = {i: v for i, v in enumerate(z["expected"])}
    return z, clr, sv_file, spec_file, expected


@pytest.fixture(scope="module")
def pixels(case):
    return PixelCooler.from_cooler(case[1])


@pytest.mark.parametrize("source", ["cooler", "pixels"])
def test_every_sv_matches_the_reference(case, pixels, source):
    from neoloopfinder_b200.callers import Fusion
    from neoloopfinder_b200.util import load_translocation_fil
    z, clr, sv_file, _spec, expected = case
    src = clr if source == "cooler" else pixels
    svs = load_translocation_fil(sv_file, clr.binsize, z["min_intra"])
    assert len(svs) == len(z["per_sv"])
    for (c1, p1, s1, c2, p2, s2, note), want in zip(svs, z["per_sv"]):
        fu = Fusion(src, c1, c2, p1, p2, s1, s2, note, span=z["span"], col=z["col"], protocol="insitu", trim=False,
                    expected_values=expected)
        if source == "pixels":
            assert fu._asm is not None and fu._fusion_matrix is None       # built on the device
        fu.detect_bounds()
        fu.allele_slope()
        fu.correct_heterozygosity()
        assert fu.name == want["name"] and [int(v) for v in fu.k_p] == want["k_p"]
        assert (int(fu.up_i), int(fu.down_i)) == (want["up_i"], want["down_i"]), fu.name
        for k in SCALARS:
            assert float(getattr(fu, k)) == want[k], (fu.name, k)
        assert hashlib.sha256(np.ascontiguousarray(fu.hcm).tobytes()).hexdigest() == want["hcm_sha"], fu.name


@pytest.mark.parametrize("source", ["cooler", "pixels"])
def test_assemble_matches_the_reference(case, pixels, source, tmp_path):
    from neoloopfinder_b200.assembly import assembleSV
    z, clr, sv_file, _spec, expected = case
    src = clr if source == "cooler" else pixels
    work = assembleSV(src, sv_file, span=z["span"], col=z["col"], n_jobs=1, protocol="insitu", minIntra=z["min_intra"],
                      expected_values=expected)
    assert [[list(r[0]), list(r[1]), list(r[2])] for r in work.queue] == z["queue"]
    work.build_graph()
    edges = sorted([[repr(a), repr(b), float(w)] for a, b, w in work.graph.edges(data="weight")])
    assert edges == z["edges"]
    work.find_complexSV()
    assert [work._change_format(p) for p in work.alleles] == z["alleles"] and len(z["alleles"]) >= 1
    out = str(tmp_path / "assemblies.txt")
    work.output_all(out)
    assert open(out).read() == z["output"]


def test_command_line(case, tmp_path):
    """scripts/assemble-complexSVs end to end (expected prologue included), then neoloop-caller on its output."""
    from neoloopfinder_b200 import cli, cli_assemble
    z, _clr, sv_file, spec_file, _expected = case
    prefix = str(tmp_path / "run")
    cli_assemble.run(["-O", prefix, "-B", sv_file, "-H", spec_file, "-R", str(z["span"]), "--minimum-size",
                      str(z["min_intra"]), "--logFile", str(tmp_path / "a.log")])
    assert open(prefix + ".assemblies.txt").read() == z["output"]
    loops = str(tmp_path / "loops.txt")
    cli.run(["-O", loops, "-H", spec_file, "--assembly", prefix + ".assemblies.txt", "-R", str(z["span"]), "--prob", "0.5",
             "--model", os.path.join(GOLDEN, "forest_w6.joblib"), "--logFile", str(tmp_path / "n.log")])
    rows = [l.split() for l in open(loops)]
    assert len(rows) >= 1 and all(len(r) == 7 for r in rows)            # sorted loop lines with their labels
