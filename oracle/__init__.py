"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the neoloop-caller scoring path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker or as the timed CPU baseline.  The
product package ``neoloopfinder_b200`` never imports this package and fails loudly
when its CUDA library is missing.

Contents
--------
``pyport.py``      numpy/scipy/sklearn restatement of the reference path with the same
                   third-party call sites and the same loop structure as the reference
                   (this is the timed "port" CPU baseline on the GPU box, where
                   ``/root/reference`` does not exist).
``c/nlf_oracle.c`` plain-C restatement, including the third-party arithmetic the
                   reference reaches through numpy / scipy / sklearn (pairwise sums,
                   gaussian_filter, PAVA isotonic fit + interp1d, forest traversal,
                   find_peaks / peak_widths / dbscan pooling).  Built into
                   ``oracle/_ref/libnlf_oracle.so`` by ``oracle/build.py``.
``coracle.py``     ctypes wrapper over that library.
``run_reference.py``, ``gen_golden.py``
                   import the real reference from ``/root/reference`` (build container
                   only), run it stage by stage on the Cooler shim and write the golden
                   fixtures under ``tests/golden/``.

Parity pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, executed in the build
container with numpy 2.3.5 / scipy 1.18.1 / scikit-learn 1.9.0 and committed as
``tests/golden/*.npz`` together with the generating script.
"""
