"""Builds the CUDA library in-tree (``fcvm_workbench_b200/libfcvm_b200.so``) for sm_100a."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfcvm_b200.so")


def build(force: bool = False, verbose: bool = False) -> str:
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", csrc, "clean"], stdout=subprocess.DEVNULL)
    jobs = str(min(8, os.cpu_count() or 1))
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", csrc, "-j", jobs], stdout=out)
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError("build finished without producing " + LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose=True))
