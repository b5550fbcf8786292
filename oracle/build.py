"""Build the C oracle into oracle/_ref/libnlf_oracle.so.  TEST INFRASTRUCTURE.

gcc -O2 -ffp-contract=off : the reference arithmetic (numpy / scipy / sklearn wheels built for
baseline x86-64) never fuses multiply-add, so the restatement must not either.
The reference itself is pure Python (no C sources to compile), so ``oracle/_ref`` only ever
holds this restatement; kind = "port" in bench.py's cpu_baseline.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "nlf_oracle.c")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libnlf_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-fvisibility=hidden", "-fopenmp", "-Wall", "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
