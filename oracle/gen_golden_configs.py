"""Golden fixtures for the BASELINE.json configurations, written by the REAL reference.  TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden_configs            (about 8 minutes of one core)

cfg1  (configs[0]) the five SVs of demo/K562-test-SVs.txt as ``C#`` assembly lines on a synthetic hg38-sized
      10 kb map, 2 Mb window: the configuration on which the reference's CPU path runs as it is.
cfg3s (configs[2] shape) one complex assembly of the bench workload (5 kb, 3 Mb span, 4 blocks, 1 382 bins): above
      1 024 bins the dense upload takes the band-trimmed path, and every stage K5-K10 runs at the metric's sizes.

Every assembly is stored machine-independently: the RAW integer counts of its rearranged matrix (upper triangle,
int16) and the balancing weight of every assembly bin.  The tests rebuild a pixel table (cooler's layout) from
them, so the drop-in ``pipeline()`` runs end to end -- matrix assembly, normalisation, regressions, scoring,
pooling, coordinates, flexible-assembly filter -- on exactly the numbers the reference saw.  Outputs stored: the
reference's final dictionary (scripts/neoloop-caller:185-218), the scored pixel list and probabilities (sha256
plus a strided sample), slopes, index map and the called loops.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from neoloopfinder_b200.synth import analytic_expected, make_workload  # noqa: E402
from oracle.run_reference import import_reference, run_stages  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL = os.path.join(GOLDEN, "forest_w6.joblib")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def assembly_payload(prefix, clr, ID, line, expected, span, prob):
    rc, ra, ru, cli = import_reference()
    t0 = time.time()
    st = run_stages(clr, line, expected, MODEL, span, balance="sweight", prob=prob, min_count=2, assembly_id=ID)
    # raw counts of the same rearranged matrix (the reference's own fetches with balance switched off)
    raw = ra.complexSV(clr, line, span=span, col=False, protocol="insitu", expected_values=expected, flexible=False)
    raw.reorganize_matrix()
    R = raw.fusion_matrix
    assert np.array_equal(R, np.round(R)) and R.max() < 32767 and R.min() >= 0
    # per-bin weight in assembly order, NaN -> 0 (assembly.py:394-395); check that counts x weights give back the
    # reference's balanced matrix bit for bit (cooler's order: (count * w_row) * w_col)
    w = st["bias"]
    again = np.triu((R * w[:, None]) * w[None, :])
    assert np.array_equal(again, st["fusion_matrix"]), "raw counts x weights do not reproduce the fusion matrix"
    p = {}
    p[prefix + "line"] = np.array(line)
    p[prefix + "raw"] = R.astype(np.int16)
    p[prefix + "w"] = w
    p[prefix + "block_index"] = st["block_index"]
    p[prefix + "valid"] = np.bool_(st["valid"])
    if st["valid"]:
        p[prefix + "slopes"] = st["slopes"]
        p[prefix + "index_map"] = st["index_map"]
        p[prefix + "chains"] = st["chains"]
        p[prefix + "n_cand"] = np.int64(len(st["ridx"]))
        p[prefix + "n_scored"] = np.int64(len(st["probas"]))
        p[prefix + "clist_sha"] = np.array(sha(st["clist"].astype(np.int32)))
        p[prefix + "probas_sha"] = np.array(sha(st["probas"]))
        p[prefix + "probas_sample"] = st["probas"][::97]
        p[prefix + "n_donuts"] = np.int64(len(st["donut_val"]))
        p[prefix + "loop_index"] = np.array([(i, j) for i, j, _p in st["loop_index"]], dtype=np.int32).reshape(-1, 2)
        p[prefix + "loop_prob"] = np.array([q for _i, _j, q in st["loop_index"]], dtype=np.float64)
        fd = st["final_dict"] or {}
        keys = sorted(fd)
        p[prefix + "final_keys"] = np.array(["|".join(map(str, k)) for k in keys])
        p[prefix + "final_vals"] = np.array(["|".join(map(str, fd[k])) for k in keys])
    print("%s%s: n=%d scored=%d donuts=%d loops=%d final=%d  (reference %.0f s)" % (
        prefix, ID, R.shape[0], len(st.get("probas", [])), len(st.get("donut_val", [])), len(st.get("loop_index", [])),
        len(st.get("final_dict") or {}), time.time() - t0), flush=True)
    return p


def dump(name, clr, lines, span, prob, expected):
    exp_arr = np.array([expected[i] for i in sorted(expected)], dtype=np.float64)
    payload = dict(res=np.int64(clr.binsize), span=np.int64(span), prob=np.float64(prob), expected=exp_arr,
                   cooler_spec=np.array(json.dumps(clr.spec())), ids=np.array(list(lines)),
                   model=np.array(os.path.basename(MODEL)))
    for ID, line in lines.items():
        payload.update(assembly_payload(ID + "_", clr, ID, line, expected, span, prob))
    out = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(out, **payload)
    print("%s -> %d bytes" % (out, os.path.getsize(out)), flush=True)


def main():
    which = sys.argv[1:] or ["cfg1", "cfg3s"]
    if "cfg1" in which:
        clr, lines = make_workload("demo", res=10000, seed=11, span=2_000_000)
        dump("cfg1_demo_10k", clr, lines, 2_000_000, 0.5, analytic_expected(clr))
    if "cfg3s" in which:
        clr, lines = make_workload("complex", res=5000, seed=100, n_assemblies=192, span=3_000_000)
        dump("cfg3_shape_5k", clr, {"A154": lines["A154"]}, 3_000_000, 0.5, analytic_expected(clr))


if __name__ == "__main__":
    main()
