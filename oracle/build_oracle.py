"""TEST INFRASTRUCTURE ONLY -- builds oracle/_build/libradsearch_oracle.so from oracle/radsearch_oracle.c and
oracle/maps_oracle.c with gcc.

`python oracle/build_oracle.py` (also called by __graft_entry__.build() and lazily by oracle/c_oracle.py).
-ffp-contract=off: the restatement must not fuse a*b+c, numpy and CPython do not.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libradsearch_oracle.so")
SRC = os.path.join(HERE, "radsearch_oracle.c")
SRC2 = os.path.join(HERE, "maps_oracle.c")
HDR = os.path.join(HERE, "radsearch_oracle.h")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(SRC2), os.path.getmtime(HDR)):
        return LIB
    cmd = ["gcc", "-std=c99", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-Wall", "-Wextra",
           "-o", LIB, SRC, SRC2, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
