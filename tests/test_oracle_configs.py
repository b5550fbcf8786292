"""The C oracle on the configuration fixtures written by the real reference (oracle/gen_golden_configs.py):
cfg1 (demo SVs, 10 kb / 2 Mb) and the cfg3-shaped 1 382-bin assembly at 5 kb.  CPU only, bit for bit."""
import os

import numpy as np
import pytest

from oracle import coracle as co
from tests.helpers import GOLDEN, scale_plan_from_slopes, sha


@pytest.mark.parametrize("name,ids", [("cfg1_demo_10k", ["C0", "C3"]), ("cfg3_shape_5k", ["A154"])])
def test_oracle_equals_reference_on_config_fixtures(name, ids, models):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    res, prob = int(z["res"]), float(z["prob"])
    arrays = co.sklearn_forest_arrays(models(str(z["model"])))
    for ID in ids:
        p = ID + "_"
        w = z[p + "w"]
        M = np.triu((z[p + "raw"].astype(np.float64) * w[:, None]) * w[None, :])      # cooler's order, as the generator checked
        op, val = scale_plan_from_slopes(z[p + "slopes"])
        hcm = co.scale_blocks(M, z[p + "block_index"], op, val)
        kept = co.remove_gaps_index(hcm)
        assert np.array_equal(kept, z[p + "index_map"])
        G = hcm[np.ix_(kept, kept)]
        cl, pr, _ = co.score_matrix(G, z["expected"], arrays, 6)
        assert len(pr) == int(z[p + "n_scored"])
        assert sha(cl.astype(np.int32)) == str(z[p + "clist_sha"]) and sha(pr) == str(z[p + "probas_sha"])
        keep = pr >= prob
        ij = cl[keep]
        v = np.array([G[i, j] for i, j in ij])
        chains, im = z[p + "chains"], z[p + "index_map"]
        loops = co.local_clustering(ij, v, res, 2, 15000, chains[im[ij[:, 0]]], chains[im[ij[:, 1]]])
        assert sorted(loops) == [tuple(x) for x in z[p + "loop_index"]]
