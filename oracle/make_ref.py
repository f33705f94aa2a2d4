"""TEST INFRASTRUCTURE -- puts the UNMODIFIED reference where the GPU box can run it.

    python oracle/make_ref.py

copies ``source code/fcVM.py``, ``dummyVM.py`` and ``fcVM.ini`` byte for byte from ``/root/reference``
(present only in the build container) into ``oracle/_ref/`` -- git-ignored, so no reference source enters
the history, but not gpurun-ignored, so the copy travels to the GPU box with the built libraries.
``oracle/ref_harness.py`` loads the reference from there when ``/root/reference`` is absent, and
``bench.py --impl reference`` then times the reference's own numba routines and ``calcDisp`` on the box's host
cores (CHOLMOD replaced by SuperLU, see ref_harness) instead of the C port.  ``__graft_entry__.build()`` calls
this.
"""
from __future__ import annotations

import filecmp
import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("FCVM_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(_HERE, "_ref")
FILES = (os.path.join("source code", "fcVM.py"), "dummyVM.py", "fcVM.ini")


def make_ref(verbose: bool = True) -> bool:
    """Returns True when ``oracle/_ref`` holds the reference files afterwards."""
    if not os.path.isfile(os.path.join(SRC_ROOT, FILES[0])):
        have = os.path.isfile(os.path.join(DST_ROOT, FILES[0]))
        if verbose:
            print(f"make_ref: {SRC_ROOT} not present; " + ("keeping the existing copy" if have else "no copy made"))
        return have
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST_ROOT, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    if verbose:
        print(f"make_ref: unmodified reference copied to {DST_ROOT}")
    return True


if __name__ == "__main__":
    make_ref()
