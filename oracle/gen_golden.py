"""Generate the golden fixtures under tests/golden/ from the REAL reference.  TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden

The reference ships no golden vectors, so parity is pinned on the
outputs of the reference itself, executed here stage by stage (oracle/run_reference.py)
with numpy 2.3.5 / scipy 1.18.1 / scikit-learn 1.9.0.  Each case stores the *inputs* of
the path as arrays (the assembled fusion matrix, bias, block index, expected) -- not only
the procedural map spec -- so the fixtures do not depend on libm/SIMD details of the
machine that replays them.

Fixtures written
----------------
forest_w6.joblib / forest_w5.joblib   synthetic RandomForest models (169 / 121 features),
                                      trained on windows produced by the reference's own
                                      ``Peakachu.getwindow`` (labels = injected loop cells
                                      plus label noise so that trees grow deep).
case_*.npz                            per-stage inputs/outputs of pipeline() for one assembly.
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from neoloopfinder_b200.coolshim import SyntheticCooler  # noqa: E402
from neoloopfinder_b200.synth import SMALL_CHROMSIZES  # noqa: E402
from oracle.run_reference import import_reference, run_stages  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def train_forest(res, w, n_estimators, n_neg, noise, seed, path):
    """Features come from the reference's getwindow on a training region of a map with a
    different seed than any test case."""
    import joblib
    from sklearn.ensemble import RandomForestClassifier
    rc, ra, ru, cli = import_reference()
    clr = SyntheticCooler(binsize=res, chromsizes=SMALL_CHROMSIZES, seed=1000 + seed,
                          amplitude=255.0 * res / 10000.0)
    exp = ru.calculate_expected(clr, list(SMALL_CHROMSIZES), maxdis=5000000, balance="sweight", nproc=1)
    lo_bp, hi_bp = 1_000_000, 1_000_000 + 420 * res
    M = clr.matrix(balance="sweight").fetch(("chr1", lo_bp, hi_bp))
    M[np.isnan(M)] = 0
    M = np.triu(M)
    core = rc.Peakachu(M, upper=400 * res, res=res)
    core.load_expected(exp)
    coords = list(zip(core.ridx, core.cidx))
    fea, clist = core.getwindow(coords, w)
    a, b = clr._chrom_loops("chr1")
    lo = lo_bp // res
    cells = set()
    for aa, bb in zip(a, b):
        for da in (-1, 0, 1):
            for db in (-1, 0, 1):
                cells.add((int(aa) - lo + da, int(bb) - lo + db))
    y = np.array([(int(i), int(j)) in cells for i, j in clist])
    rng = np.random.default_rng(seed)
    pos = np.where(y)[0]
    neg = np.where(~y)[0]
    sel = np.r_[pos, rng.choice(neg, min(n_neg, neg.size), replace=False)]
    X = fea[sel]
    Y = y[sel].astype(int)
    Y = np.where(rng.random(Y.size) < noise, 1 - Y, Y)
    rf = RandomForestClassifier(n_estimators=n_estimators, max_depth=20, random_state=42, n_jobs=1).fit(X, Y)
    joblib.dump(rf, path, compress=3)
    nodes = sum(e.tree_.node_count for e in rf.estimators_)
    print(f"forest {os.path.basename(path)}: {X.shape[0]} samples, {nodes} nodes, {os.path.getsize(path)} bytes")
    return rf


def dump_case(name, clr, line, chroms_for_expected, span, model_path, prob, balance="sweight",
              min_count=2, fea_stride=16):
    rc, ra, ru, cli = import_reference()
    expected = ru.calculate_expected(clr, chroms_for_expected, maxdis=5000000, balance=balance, nproc=1)
    t0 = time.time()
    st = run_stages(clr, line, expected, model_path, span, balance=balance, prob=prob, min_count=min_count)
    dt = time.time() - t0
    exp_arr = np.array([expected[i] for i in sorted(expected)], dtype=np.float64)
    payload = dict(
        res=np.int64(clr.binsize), span=np.int64(span), prob=np.float64(prob), min_count=np.int64(min_count),
        line=np.array(line), balance=np.array(str(balance)), model=np.array(os.path.basename(model_path)),
        cooler_spec=np.array(__import__("json").dumps(clr.spec())),
        expected=exp_arr, valid=np.bool_(st["valid"]),
        fusion_matrix=st["fusion_matrix"], bias=st["bias"], block_index=st["block_index"],
        orients=np.array(st["orients"]),
    )
    if st["valid"]:
        nb = st["block_index"].shape[0]
        # local expected per block pair at maxdis = 2 Mb (superset of the shorter ones)
        for (a, b), d in st["exps"].items():
            keys = np.array(sorted(d), dtype=np.int64)
            payload[f"exps_idx_{a}_{b}"] = keys
            payload[f"exps_val_{a}_{b}"] = np.array([d[k] for k in keys], dtype=np.float64)
        fea = st["fea"]
        fea32 = fea.astype(np.float32)
        payload.update(
            slopes=st["slopes"], hcm_sha=np.array(sha(st["hcm"])),
            gap_free_sha=np.array(sha(st["gap_free"])), index_map=st["index_map"], chains=st["chains"],
            w=np.int64(st["w"]), ridx=st["ridx"].astype(np.int32), cidx=st["cidx"].astype(np.int32),
            clist=st["clist"].astype(np.int32), fea_sha=np.array(sha(fea)), fea32_sha=np.array(sha(fea32)),
            fea_rows=np.arange(0, fea.shape[0], fea_stride, dtype=np.int64),
            fea_sample=fea[::fea_stride], probas=st["probas"],
            donut_ij=st["donut_ij"].astype(np.int32), donut_val=st["donut_val"],
            loop_index=np.array([(i, j) for i, j, p in st["loop_index"]], dtype=np.int32).reshape(-1, 2),
            loop_prob=np.array([p for i, j, p in st["loop_index"]], dtype=np.float64),
        )
        fd = st["final_dict"] or {}
        keys = sorted(fd)
        payload["final_keys"] = np.array(["|".join(map(str, k)) for k in keys])
        payload["final_vals"] = np.array(["|".join(map(str, fd[k])) for k in keys])
    out = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(out, **payload)
    n = st["fusion_matrix"].shape[0]
    print(f"{name}: n={n} valid={st['valid']} cand={len(st.get('ridx', []))} scored={len(st.get('probas', []))} "
          f"donuts={len(st.get('donut_val', []))} loops={len(st.get('loop_index', []))} "
          f"final={len(st.get('final_dict') or {})}  ref {dt:.1f}s  -> {os.path.getsize(out)} bytes")


def dump_pool_cases(path, n_cases=60, seed=99):
    """Golden vectors for the pooling stage alone: the reference's own
    ``Peakachu.local_clustering`` (neoloop/callers.py:668-831) on synthetic Donut sets."""
    rc, ra, ru, cli = import_reference()
    rng = np.random.default_rng(seed)
    payload = {}
    for k in range(n_cases):
        res = int(rng.choice([5000, 10000, 25000]))
        n = int(rng.integers(60, 400))
        n_clusters = int(rng.integers(1, 25))
        pts = {}
        for _ in range(n_clusters):
            ci = int(rng.integers(0, n - 10))
            cj = int(rng.integers(ci + 8, min(n, ci + 200) + 8))
            m = int(rng.integers(1, 30))
            spread = float(rng.choice([0.7, 1.5, 3.0, 6.0]))
            for _ in range(m):
                i = int(round(ci + rng.normal(0, spread)))
                j = int(round(cj + rng.normal(0, spread)))
                if 0 <= i < j < n + 210:
                    # quantised values so that exact ties occur
                    pts[(i, j)] = float(rng.integers(1, 40)) * 0.125 if rng.random() < 0.3 else float(rng.random() * 5)
        for _ in range(int(rng.integers(0, 15))):
            i = int(rng.integers(0, n)); j = int(rng.integers(i + 1, n + 200))
            pts[(i, j)] = float(rng.random() * 5)
        if not pts:
            pts[(3, 20)] = 1.0
        nmax = max(max(p) for p in pts) + 1
        min_count = int(rng.choice([1, 2, 2, 3]))
        use_chains = rng.random() < 0.7
        core = object.__new__(rc.Peakachu)
        core.r = res
        ij = np.array(list(pts), dtype=np.int64)
        val = np.array([pts[tuple(p)] for p in ij], dtype=np.float64)
        if use_chains:
            cuts = np.sort(rng.choice(np.arange(1, nmax), size=min(nmax - 1, int(rng.integers(1, 4))), replace=False))
            chain_of = np.searchsorted(cuts, np.arange(nmax), side="right")
            chains = {int(i): int(c) for i, c in enumerate(chain_of)}
            got = core.local_clustering(dict(pts), min_count=min_count, index_map=None, chains=chains)
        else:
            chain_of = np.zeros(nmax, dtype=np.int64)
            got = core.local_clustering(dict(pts), min_count=min_count)
        got = np.array(sorted((int(a), int(b)) for a, b in got), dtype=np.int64).reshape(-1, 2)
        payload[f"c{k}_ij"] = ij.astype(np.int32)
        payload[f"c{k}_val"] = val
        payload[f"c{k}_chain"] = chain_of.astype(np.int32)
        payload[f"c{k}_par"] = np.array([res, min_count, 15000], dtype=np.int64)
        payload[f"c{k}_out"] = got.astype(np.int32)
    payload["n_cases"] = np.int64(n_cases)
    np.savez_compressed(path, **payload)
    print(f"pool cases: {n_cases} -> {os.path.getsize(path)} bytes")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    dump_pool_cases(os.path.join(GOLDEN, "pool_cases.npz"))
    f6 = os.path.join(GOLDEN, "forest_w6.joblib")
    f5 = os.path.join(GOLDEN, "forest_w5.joblib")
    if not os.path.exists(f6):
        train_forest(10000, 6, 100, 3500, 0.06, 7, f6)
    if not os.path.exists(f5):
        train_forest(25000, 5, 20, 1500, 0.05, 8, f5)

    cs = SMALL_CHROMSIZES
    chroms = list(cs)
    # 1. translocation +-, 10 kb
    sv = ("chr1", 5_003_000, "+", "chr2", 4_000_500, "-")
    clr = SyntheticCooler(binsize=10000, chromsizes=cs, seed=1, svs=[sv])
    dump_case("case_transloc_10k", clr, "translocation,1,5003000,+,2,4000500,-\t1,4200000\t2,4800000",
              chroms, 800_000, f6, prob=0.15)
    # 2. inversion --, 10 kb (both blocks flipped relative to each other), tighter threshold
    sv = ("chr3", 2_000_000, "-", "chr3", 4_500_000, "-")
    clr = SyntheticCooler(binsize=10000, chromsizes=cs, seed=2, svs=[sv])
    dump_case("case_inversion_10k", clr, "inversion,3,2000000,-,3,4500000,-\t3,2700000\t3,5200000",
              chroms, 700_000, f6, prob=0.6)
    # 3. three-SV chain with a reversed middle block, 10 kb
    svs = [("chr1", 3_000_000, "+", "chr2", 6_000_000, "+"), ("chr2", 5_600_000, "-", "chr4", 2_000_000, "-")]
    clr = SyntheticCooler(binsize=10000, chromsizes=cs, seed=3, svs=svs)
    dump_case("case_complex_10k", clr,
              "translocation,1,3000000,+,2,6000000,+\ttranslocation,2,5600000,-,4,2000000,-\t1,2500000\t4,2500000",
              chroms, 500_000, f6, prob=0.5)
    # 4. deletion at 5 kb (expected array of 1001 entries)
    sv = ("chr2", 3_000_000, "+", "chr2", 5_000_000, "-")
    clr = SyntheticCooler(binsize=5000, chromsizes=cs, seed=4, svs=[sv], amplitude=127.5)
    dump_case("case_deletion_5k", clr, "deletion,2,3000000,+,2,5000000,-\t2,2600000\t2,5400000",
              chroms, 450_000, f6, prob=0.5)
    # 5. 25 kb, w=5 model; n > 213 so windows with d+2w >= 201 take the un-normalised branch
    sv = ("chr1", 6_000_000, "+", "chr3", 3_000_000, "-")
    clr = SyntheticCooler(binsize=25000, chromsizes=cs, seed=5, svs=[sv], amplitude=637.5)
    dump_case("case_transloc_25k", clr, "translocation,1,6000000,+,3,3000000,-\t1,3000000\t3,6000000",
              chroms, 3_000_000, f5, prob=0.5, fea_stride=32)
    # 6. invalid assembly (overlapping blocks => len(Map) != n; pipeline returns None)
    sv = ("chr1", 5_000_000, "-", "chr1", 5_300_000, "+")
    clr = SyntheticCooler(binsize=10000, chromsizes=cs, seed=6, svs=[])
    dump_case("case_invalid_10k", clr, "duplication,1,5000000,-,1,5300000,+\t1,5800000\t1,4500000",
              chroms, 800_000, f6, prob=0.5)


if __name__ == "__main__":
    main()
