#!/usr/bin/env python
"""Compile oracle/gsr_oracle.c (CPU restatement, test infrastructure) with gcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gsr_oracle.c")
LIB = os.path.join(HERE, "liboracle.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                    "-o", LIB, SRC, "-lm"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
