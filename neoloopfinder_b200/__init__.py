"""neoloopfinder_b200 -- the neoloop-caller scoring path of NeoLoopFinder on NVIDIA B200.

Hand-written CUDA (sm_100a) behind a C ABI (``include/nlf_b200.h``, ``libnlf_b200.so``) with a
Python host layer that keeps the reference's class and command-line interface:

    neoloopfinder_b200.assembly.complexSV      <- neoloop.assembly.complexSV
    neoloopfinder_b200.callers.Peakachu        <- neoloop.callers.Peakachu
    neoloopfinder_b200.callers.combine_annotations
    neoloopfinder_b200.util.calculate_expected
    neoloopfinder_b200.cli.run / pipeline      <- scripts/neoloop-caller

There is no CPU fallback: importing the package is cheap, the first GPU call loads the library
and raises if it or a B200 is missing.
"""
__version__ = "0.5.0+b200.1"
