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
