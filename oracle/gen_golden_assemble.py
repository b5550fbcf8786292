"""Golden fixture for row f-3 (assemble-complexSVs inner loop) written by the REAL reference.
TEST INFRASTRUCTURE; build container only (needs /root/reference):

    python -m oracle.gen_golden_assemble            -> tests/golden/assemble_10k.json

The reference's own ``assembleSV`` (filterSV on every SV, build_graph, find_complexSV, output_all) runs
unmodified on a procedural map; the fixture keeps the map's spec, the SV list, the expected table and every
result the GPU tests compare: per-SV bounds / slopes / R^2, the surviving SVs, graph edges, alleles and the
assembly file text."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from neoloopfinder_b200.coolshim import SyntheticCooler          # noqa: E402  (input generator only)
from oracle.run_reference import import_reference                 # noqa: E402

SIZES = {"chr1": 12_000_000, "chr2": 9_000_000, "chr3": 7_000_000, "chr4": 6_000_000}
SV_LINES = [
    "chr1 chr2 +- 6000000 4000000 translocation",
    "chr2 chr3 +- 4150000 2000000 translocation",           # 150 kb after the first one: the two can chain
    "chr3 chr4 +- 2500000 3000000 translocation",
    "chr1 chr1 +- 2000000 4000000 deletion",
    "chr2 chr2 ++ 6000000 7500000 inversion",
    "chr4 chr4 -- 1000000 2200000 inversion",
    "chr3 chr1 -+ 5000000 9000000 translocation",
    "chr4 chr2 ++ 5000000 1000000 translocation",           # no induced contacts in the map: must be filtered out
]
# contacts that only exist on the chained allele chr1 -> chr2[4.00-4.15 Mb] -> chr3 (map only, not an input SV)
MAP_ONLY_SVS = [("chr1", 6000000, "+", "chr3", 2000000, "-", 0.35)]
SPAN, RES, COL = 800000, 10000, "sweight"


def main():
    rc, ra, ru, _cli = import_reference()
    svs = []
    for line in SV_LINES[:-1]:
        c1, c2, st, p1, p2, _note = line.split()
        svs.append((c1, int(p1), st[0], c2, int(p2), st[1]))
    clr = SyntheticCooler(binsize=RES, chromsizes=SIZES, seed=31, svs=svs + MAP_ONLY_SVS)
    chroms = list(SIZES)
    expected = ru.calculate_expected(clr, chroms, maxdis=5000000, balance=COL, nproc=1)
    tmp = tempfile.mkdtemp()
    sv_file = os.path.join(tmp, "svs.txt")
    open(sv_file, "w").write("\n".join(SV_LINES) + "\n")
    per_sv = []
    for c1, p1, s1, c2, p2, s2, note in ru.load_translocation_fil(sv_file, RES, 500000):
        fu = rc.Fusion(clr, c1, c2, p1, p2, s1, s2, note, span=SPAN, col=COL, protocol="insitu", trim=False,
                       expected_values=expected)
        fu.detect_bounds()
        fu.allele_slope()
        fu.correct_heterozygosity()
        per_sv.append({"name": fu.name, "k_p": [int(v) for v in fu.k_p], "up_i": int(fu.up_i), "down_i": int(fu.down_i),
                       "u_g_r": float(fu.u_g_r), "u_g_s": float(fu.u_g_s), "d_g_r": float(fu.d_g_r),
                       "d_g_s": float(fu.d_g_s), "m_g_r": float(fu.m_g_r), "m_g_s": float(fu.m_g_s),
                       "u_s": float(fu.u_s), "d_s": float(fu.d_s),
                       "hcm_sum": float(fu.hcm.sum()), "hcm_sha": __import__("hashlib").sha256(np.ascontiguousarray(fu.hcm).tobytes()).hexdigest()})
    work = ra.assembleSV(clr, sv_file, span=SPAN, col=COL, n_jobs=1, protocol="insitu", minIntra=500000,
                         expected_values=expected)
    work.build_graph()
    work.find_complexSV()
    out = os.path.join(tmp, "assemblies.txt")
    work.output_all(out)
    fixture = {
        "cooler_spec": clr.spec(), "sv_lines": SV_LINES, "span": SPAN, "col": COL, "min_intra": 500000,
        "expected": [float(expected[i]) for i in sorted(expected)],
        "per_sv": per_sv,
        "queue": [[list(r[0]), list(r[1]), list(r[2])] for r in work.queue],
        "edges": sorted([[repr(a), repr(b), float(w)] for a, b, w in work.graph.edges(data="weight")]),
        "alleles": [work._change_format(p) for p in work.alleles],
        "output": open(out).read(),
        "versions": {"numpy": np.__version__, "sklearn": __import__("sklearn").__version__,
                     "scipy": __import__("scipy").__version__, "networkx": __import__("networkx").__version__},
    }
    path = os.path.join(ROOT, "tests", "golden", "assemble_10k.json")
    json.dump(fixture, open(path, "w"), indent=1)
    print("wrote", path)
    print("kept SVs:", len(work.queue), "edges:", work.graph.number_of_edges(), "alleles:", len(work.alleles))
    print(fixture["output"])


if __name__ == "__main__":
    main()
