"""Shared helpers for the parity tests (host-side restatement of tiny glue only)."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_case(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def scale_plan_from_slopes(sl):
    """op/val per block pair from the (rscore, slope) table, complexSV.correct_heterozygosity
    (neoloop/assembly.py:536-549).  op: 0 none, 1 divide, 2 multiply."""
    nb = sl.shape[0]
    op = np.zeros((nb, nb), np.int32)
    val = np.ones((nb, nb))
    for a in range(nb):
        for b in range(a, nb):
            ms = max(sl[a, a, 1], sl[b, b, 1])
            if sl[a, b, 0] > 0.6 and sl[a, b, 1] > 0 and ms > 0:
                if a == b:
                    op[a, b], val[a, b] = 1, sl[a, b, 1]
                else:
                    op[a, b], val[a, b] = 2, min(1 / sl[a, b, 1], 5 / ms)
    return op, val


def final_dict_of(z, prefix):
    out = {}
    for k, v in zip(z[prefix + "final_keys"], z[prefix + "final_vals"]):
        c1, s1, e1, c2, s2, e2 = str(k).split("|")
        ID, gdis, neo = str(v).split("|")
        out[(c1, int(s1), int(e1), c2, int(s2), int(e2))] = (ID, int(gdis), int(neo))
    return out


def pixel_cooler_from_fixture(z):
    """Rebuild the sample of a configuration fixture (oracle/gen_golden_configs.py) as a pixel table in cooler's
    layout: the raw counts of every assembly's rearranged matrix go back to genome coordinates through the
    assembly's own bin map, the stored per-bin weights fill both weight columns (NaN outside the assemblies).
    Returns (PixelCooler, {ID: line})."""
    import json

    from neoloopfinder_b200.assembly import complexSV
    from neoloopfinder_b200.coolshim import PixelCooler, SyntheticCooler
    geo = SyntheticCooler.from_spec(json.loads(str(z["cooler_spec"])))       # geometry only: no pixel is generated
    res = geo.binsize
    names = list(geo.chromnames)
    nb = [geo._nbins(c) for c in names]
    offs = dict(zip(names, np.r_[0, np.cumsum(nb)[:-1]].tolist()))
    total = int(sum(nb))
    w = np.full(total, np.nan)
    keys, vals, lines = [], [], {}
    for ID in (str(x) for x in z["ids"]):
        line = str(z[ID + "_line"])
        lines[ID] = line
        fu = complexSV(geo, line, span=int(z["span"]), col="sweight", expected_values={}, flexible=False)
        fu.reorganize_matrix(map_only=True)
        n = fu._n_bins
        g = np.array([offs[fu.reverse[k][0]] + fu.reverse[k][1] // res for k in range(n)], dtype=np.int64)
        w[g] = z[ID + "_w"]
        R = z[ID + "_raw"]
        i, j = np.nonzero(R)
        keys.append(np.minimum(g[i], g[j]) * total + np.maximum(g[i], g[j]))
        vals.append(R[i, j].astype(np.int32))
    key = np.concatenate(keys)
    val = np.concatenate(vals)
    key, first = np.unique(key, return_index=True)
    val = val[first]
    bin1 = key // total
    indptr = np.zeros(total + 1, dtype=np.int64)
    np.cumsum(np.bincount(bin1, minlength=total), out=indptr[1:])
    sizes = {c: int(geo._chromsizes[c]) for c in names}
    return PixelCooler(res, sizes, indptr, (key - bin1 * total).astype(np.int32), val, {"weight": w, "sweight": w.copy()}), lines


def pool_child(n, stop):
    """Body of a spawned worker process that farms work out to its own loky pool (tests/test_host_boundary.py)."""
    from neoloopfinder_b200.sharding import estimate_pixels
    from neoloopfinder_b200.util import run_parallel, shutdown_workers
    try:
        return sum(run_parallel(2, estimate_pixels, range(400, 400 + n)))
    finally:
        if stop:
            shutdown_workers()
