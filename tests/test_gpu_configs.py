"""BASELINE.json configurations through the drop-in ``pipeline()`` against fixtures written by the REAL reference
(oracle/gen_golden_configs.py): cfg1 = the five demo SVs at 10 kb / 2 Mb, cfg3 shape = a 4-block, 1 382-bin complex
assembly at 5 kb / 3 Mb.  The sample is rebuilt from the fixture as a pixel table, so matrix assembly, normalisation,
regressions, scoring, pooling, coordinates and the flexible-assembly filter all run -- on the device from resident
pixels, and once more through the host-matrix constructor (dense upload; band-trimmed above 1 024 bins)."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN, final_dict_of, pixel_cooler_from_fixture, sha

pytestmark = pytest.mark.gpu


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    clr, lines = pixel_cooler_from_fixture(z)
    expected = {i: float(v) for i, v in enumerate(z["expected"])}
    return z, clr, lines, expected


@pytest.mark.parametrize("name", ["cfg1_demo_10k", "cfg3_shape_5k"])
def test_pipeline_equals_the_reference(name):
    from neoloopfinder_b200.cli import call_assemblies, pipeline
    z, clr, lines, expected = _load(name)
    res, span, prob = int(z["res"]), int(z["span"]), float(z["prob"])
    models = {res: os.path.join(GOLDEN, str(z["model"]))}
    stats = {}
    batch = call_assemblies(clr, list(lines), lines, span, "sweight", prob, 2, False, models[res], expected, stats=stats)
    n_scored = 0
    for ID in lines:
        want = final_dict_of(z, ID + "_")
        assert pipeline(clr, ID, lines, span, "sweight", prob, 2, False, models, expected) == want, ID
        assert batch[ID] == want, ID
        n_scored += int(z[ID + "_n_scored"])
    assert stats["scored"] == n_scored
    assert sum(len(v) for v in batch.values()) > 0


@pytest.mark.parametrize("name", ["cfg1_demo_10k", "cfg3_shape_5k"])
def test_every_stage_equals_the_reference(name):
    """The reference-shaped classes driven like scripts/neoloop-caller:185-218, stage outputs against the fixture;
    then the same gap-free matrix through the host-matrix constructor (nlf_batch_add_dense)."""
    from neoloopfinder_b200.assembly import complexSV
    from neoloopfinder_b200.callers import Peakachu
    z, clr, lines, expected = _load(name)
    res, span, prob = int(z["res"]), int(z["span"]), float(z["prob"])
    model = os.path.join(GOLDEN, str(z["model"]))
    for ID, line in lines.items():
        p = ID + "_"
        fu = complexSV(clr, line, span=span, col="sweight", expected_values=expected, flexible=False)
        fu.reorganize_matrix()
        assert len(fu.Map) == fu._n_bins
        assert [tuple(x) for x in fu.index] == [tuple(x) for x in z[p + "block_index"]]
        fu.correct_heterozygosity()
        for (a, b), (r, s) in fu.slopes_.items():
            assert r == z[p + "slopes"][a, b, 0] and s == z[p + "slopes"][a, b, 1], (ID, a, b)
        fu.remove_gaps()
        assert np.array_equal(np.array([fu.index_map[i] for i in range(len(fu.index_map))]), z[p + "index_map"])
        loops_by_route = []
        for route in ("device", "host"):
            G = fu.gap_free if route == "device" else np.ascontiguousarray(np.asarray(fu.gap_free))
            core = Peakachu(G, upper=400 * res, res=res)
            core.load_expected(expected)
            core.load_models(model)
            assert len(core.ridx) == int(z[p + "n_cand"])
            loops = sorted(core.predict(thre=prob, no_pool=False, min_count=2, index_map=fu.index_map, chains=fu.chains))
            clist, probas = core.scored_pixels()
            assert len(probas) == int(z[p + "n_scored"])
            assert sha(clist.astype(np.int32)) == str(z[p + "clist_sha"]), (ID, route)
            assert sha(probas) == str(z[p + "probas_sha"]), (ID, route)
            assert np.array_equal(probas[::97], z[p + "probas_sample"])
            want = [(int(a), int(b), float(q)) for (a, b), q in zip(z[p + "loop_index"], z[p + "loop_prob"])]
            assert loops == want, (ID, route)
            loops_by_route.append(loops)
            core.close()
        assert loops_by_route[0] == loops_by_route[1]
        if name == "cfg3_shape_5k":
            assert fu.gap_free.shape[0] > 1024 and len(fu.blocks) >= 4      # band-trimmed upload, >= 4 blocks
