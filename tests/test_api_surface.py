"""the Python surface that must not change.  Every listed class method / function exists with
the reference's parameter names and defaults (build container only: signatures are read from the reference tree)."""
import inspect

import pytest

from oracle.run_reference import reference_available

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")]

SURFACE = {
    "callers.Fusion": ["__init__", "load_expected", "get_matrix", "extract_bias", "detect_bounds", "extend_stretches", "locate",
                       "linear_regression", "allele_slope", "correct_heterozygosity", "remove_gaps"],
    "callers.Peakachu": ["__init__", "get_candidates", "load_models", "load_expected", "getwindow", "predict",
                         "local_clustering", "_local_clustering", "find_anchors", "_cluster_core"],
    "assembly.complexSV": ["__init__", "parse_input", "get_single_block", "get_bias", "get_matrix", "reorganize_matrix",
                           "correct_heterozygosity", "index_to_coordinates", "remove_gaps"],
    "assembly.assembleSV": ["__init__", "build_graph", "find_complexSV", "connect", "_containloop", "_getreverse", "_issubset",
                            "_check_overlap", "_change_format", "print", "output_all"],
}
FUNCTIONS = {"callers": ["check_increasing", "clean_inputs", "combine_annotations"],
             "assembly": ["filterSV", "filterAssembly"],
             "util": ["calculate_expected", "distance_normalize", "image_normalize", "map_coordinates", "load_translocation_fil",
                      "find_chrom_pre"]}


def _params(fn):
    return [(p.name, p.default if p.default is not inspect.Parameter.empty else "<required>")
            for p in inspect.signature(fn).parameters.values()]


def test_classes_and_functions_keep_their_signatures():
    from oracle.run_reference import import_reference
    rc, ra, ru, _cli = import_reference()
    import neoloopfinder_b200.assembly as ma
    import neoloopfinder_b200.callers as mc
    import neoloopfinder_b200.util as mu
    ref = {"callers": rc, "assembly": ra, "util": ru}
    mine = {"callers": mc, "assembly": ma, "util": mu}
    for dotted, methods in SURFACE.items():
        mod, cls = dotted.split(".")
        rcls, mcls = getattr(ref[mod], cls), getattr(mine[mod], cls)
        for name in methods:
            assert hasattr(mcls, name), (dotted, name)
            want = _params(getattr(rcls, name))
            got = _params(getattr(mcls, name))
            # optional additions are allowed after the reference's parameters (e.g. reorganize_matrix(map_only=False))
            assert got[:len(want)] == want, (dotted, name, got, want)
            assert all(d != "<required>" for _n, d in got[len(want):]), (dotted, name)
    for mod, names in FUNCTIONS.items():
        for name in names:
            want = _params(getattr(ref[mod], name))
            got = _params(getattr(mine[mod], name))
            assert got[:len(want)] == want, (mod, name, got, want)
    assert issubclass(ma.complexSV, mc.Fusion)
