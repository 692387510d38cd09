"""Compile the plain-C oracle into oracle/_build/librqae_oracle.so (test infrastructure).

The reference is pure Python (no compiled sources), so there is no ``oracle/_ref``; the
reference itself is exercised by importing it in the build container
(tests/golden/make_golden.py)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
SO = os.path.join(OUT_DIR, "librqae_oracle.so")
SRC = os.path.join(HERE, "rqae_oracle.c")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-mavx2", "-mfma", "-fopenmp",
           "-shared", "-fPIC", "-o", SO, SRC, "-lm"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
