"""Multi-GPU host scheduler (one worker process -- or, on request, one thread -- and one context per GPU, assemblies
sharded by estimated pixels, no collective): the merged result must not depend on the number of devices.
Needs two visible GPUs (``gpurun --gpus 2``); skipped on a single-GPU box."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu


def test_call_resolution_same_on_one_and_two_gpus():
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.cli import call_resolution, merge_results
    from neoloopfinder_b200.synth import analytic_expected, make_workload
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    clr, lines = make_workload("simple", res=10000, seed=321, n_assemblies=10, span=1500000)
    expected = analytic_expected(clr)
    model = os.path.join(GOLDEN, "forest_w6.joblib")
    stats1, stats2 = {}, {}
    one = call_resolution(clr, lines, 1500000, "sweight", 0.5, 2, False, model, expected, devices=[0], stats=stats1)
    two = call_resolution(clr, lines, 1500000, "sweight", 0.5, 2, False, model, expected, devices=[0, 1], stats=stats2,
                          workers="thread")
    assert one == two
    assert stats1["scored"] == stats2["scored"] > 10000
    assert len(merge_results(two)) > 0
    # the second device really worked: its context exists and launched kernels
    assert _lib.Context.get(1).launch_count() > 0


def test_resident_pixel_table_on_two_gpus():
    """Row f-2 across devices: every GPU holds its own copy of the pixel table (uploaded by the thread
    that drives it); the merged calls equal the single-GPU and the cooler-surface results."""
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.cli import call_resolution
    from neoloopfinder_b200.coolshim import PixelCooler
    from neoloopfinder_b200.synth import SMALL_CHROMSIZES, analytic_expected, make_workload
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    clr, lines = make_workload("simple", res=10000, seed=77, n_assemblies=8, chromsizes=SMALL_CHROMSIZES, span=1000000)
    px = PixelCooler.from_cooler(clr)
    expected = analytic_expected(clr)
    model = os.path.join(GOLDEN, "forest_w6.joblib")
    host = call_resolution(clr, lines, 1000000, "sweight", 0.5, 2, False, model, expected, devices=[0])
    one = call_resolution(px, lines, 1000000, "sweight", 0.5, 2, False, model, expected, devices=[0])
    two = call_resolution(px, lines, 1000000, "sweight", 0.5, 2, False, model, expected, devices=[0, 1], workers="thread")
    assert host == one == two
    assert sum(len(r) for r in two if r) > 0
    assert len({k[1] for k in px._device}) == 2          # one resident copy per device


def test_one_worker_process_per_gpu():
    """Default scheduler: a spawned worker process per GPU (own interpreter, own context, own copy of the pixel
    table); duplicated device indices collapse; results and counters equal the single-GPU run.  Same for the
    genome-wide call and the multi-sample batch."""
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.cli import call_resolution
    from neoloopfinder_b200.coolshim import PixelCooler
    from neoloopfinder_b200.genome import call_genome
    from neoloopfinder_b200.multisample import call_samples
    from neoloopfinder_b200.synth import SMALL_CHROMSIZES, analytic_expected, make_workload
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    clr, lines = make_workload("simple", res=10000, seed=77, n_assemblies=8, chromsizes=SMALL_CHROMSIZES, span=1000000)
    px = PixelCooler.from_cooler(clr)
    expected = analytic_expected(clr)
    model = os.path.join(GOLDEN, "forest_w6.joblib")
    s1, s2 = {}, {}
    one = call_resolution(px, lines, 1000000, "sweight", 0.5, 2, False, model, expected, devices=[0], stats=s1)
    two = call_resolution(px, lines, 1000000, "sweight", 0.5, 2, False, model, expected, devices=[0, 1, 1, 0], stats=s2, nproc=4)
    assert one == two and s1["scored"] == s2["scored"] > 10000
    g1 = call_genome(px, model, expected, prob=0.5, devices=[0])
    g2 = call_genome(px, model, expected, prob=0.5, devices=[0, 1])
    assert list(g1) == list(g2) and all(all(np.array_equal(a, b) for a, b in zip(g1[c], g2[c])) for c in g1)
    samples = {"A": {"clr": px, "assemblies": lines, "model": model, "expected": expected},
               "B": {"clr": clr, "assemblies": dict(list(lines.items())[:3]), "model": model, "expected": expected}}
    m1 = call_samples(samples, 1000000, "sweight", 0.5, 2, False, devices=[0])
    m2 = call_samples(samples, 1000000, "sweight", 0.5, 2, False, devices=[0, 1])
    assert m1 == m2 and m1["A"] == one
