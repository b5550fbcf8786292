"""Run the REAL reference (read from /root/reference) stage by stage.  TEST INFRASTRUCTURE.

Only usable in the build container: the GPU box has no /root/reference.  Used by
``gen_golden.py`` to write fixtures and by ``tests/test_oracle_vs_reference.py`` (skipped
when the reference is absent) to pin the oracle restatements against the reference itself.

The stages are exactly those of ``pipeline()`` (scripts/neoloop-caller:185-218); the
reference's own classes are called unmodified, this file only captures intermediates.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("NLF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "neoloop"))


def import_reference():
    """Returns (neoloop.callers, neoloop.assembly, neoloop.util, cli_module)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import neoloop.callers as rc
    import neoloop.assembly as ra
    import neoloop.util as ru
    path = os.path.join(REFERENCE_ROOT, "scripts", "neoloop-caller")
    loader = importlib.machinery.SourceFileLoader("_ref_neoloop_caller_cli", path)
    spec = importlib.util.spec_from_loader(loader.name, loader)
    cli = importlib.util.module_from_spec(spec)
    loader.exec_module(cli)
    return rc, ra, ru, cli


def run_stages(clr, line, expected, model_path, span, balance="sweight", prob=0.9, min_count=2,
               no_pool=False, assembly_id="C0"):
    """Replays pipeline() with the reference's own objects and returns every intermediate."""
    rc, ra, ru, cli = import_reference()
    out = {}
    fu = ra.complexSV(clr, line, span=span, col=balance, protocol="insitu",
                      expected_values=expected, flexible=False)
    fu.reorganize_matrix()
    out["fusion_matrix"] = fu.fusion_matrix.copy()
    out["bias"] = np.asarray(fu.bias, dtype=np.float64).copy()
    out["block_index"] = np.asarray(fu.index, dtype=np.int64)
    out["orients"] = list(fu.orients)
    out["map_size"] = len(fu.Map)
    if len(fu.Map) != fu.fusion_matrix.shape[0]:
        out["valid"] = False
        return out
    out["valid"] = True
    fu.correct_heterozygosity()
    nb = len(fu.blocks)
    rs = np.zeros((nb, nb, 2))
    for (a, b), (r, s) in fu.slopes_.items():
        rs[a, b] = (r, s)
    out["slopes"] = rs
    # local expected of the LAST maxdis (2 Mb) per pair: a prefix-closed superset of the others
    out["exps"] = {k: dict(v) for k, v in fu.exps.items()}
    out["hcm"] = fu.hcm.copy()
    out["pair_binary"] = dict(fu.pair_binary)
    fu.remove_gaps()
    out["gap_free"] = fu.gap_free.copy()
    n_gf = fu.gap_free.shape[0]
    out["index_map"] = np.array([fu.index_map[i] for i in range(n_gf)], dtype=np.int64)
    out["chains"] = np.array([fu.chains[i] for i in range(fu.fusion_matrix.shape[0])], dtype=np.int64)

    core = rc.Peakachu(fu.gap_free, upper=400 * fu.res, res=fu.res, protocol="insitu")
    core.load_expected(expected_values=expected)
    core.load_models(model_path)
    out["w"] = core.w
    out["exp_arr"] = core.exp_arr.copy()
    out["ridx"] = np.asarray(core.ridx, dtype=np.int64)
    out["cidx"] = np.asarray(core.cidx, dtype=np.int64)
    coords = [(r, c) for r, c in zip(core.ridx, core.cidx)]
    fea, clist = core.getwindow(coords, core.w)
    if len(fea) >= 2:
        out["fea"] = np.asarray(fea, dtype=np.float64)
        out["clist"] = np.asarray(clist, dtype=np.int64)
        probas = core.model.predict_proba(fea)[:, 1]
        out["probas"] = probas
        keep = probas >= prob
        ri, ci = out["clist"][keep, 0], out["clist"][keep, 1]
        out["donut_ij"] = np.stack([ri, ci], axis=1)
        out["donut_val"] = np.array([core.M[i, j] for i, j in zip(ri, ci)], dtype=np.float64)
    else:
        out["fea"] = np.zeros((0, (2 * core.w + 1) ** 2))
        out["clist"] = np.zeros((0, 2), dtype=np.int64)
        out["probas"] = np.zeros(0)
        out["donut_ij"] = np.zeros((0, 2), dtype=np.int64)
        out["donut_val"] = np.zeros(0)
    loop_index = core.predict(thre=prob, no_pool=no_pool, min_count=min_count,
                              index_map=fu.index_map, chains=fu.chains)
    out["loop_index"] = sorted((int(i), int(j), float(p)) for i, j, p in loop_index)
    out["final_dict"] = cli.pipeline(clr, assembly_id, {assembly_id: line}, span, balance, prob,
                                     min_count, no_pool, {fu.res: model_path}, expected)
    return out
