"""The C-ABI library: built, loadable, exports every symbol include/fcvm_b200.h declares.
No compute is attempted here (this suite runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fcvm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fcvm_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from fcvm_workbench_b200 import _lib, build
    if not os.path.isfile(_lib.LIB_PATH):
        build.build()
    return _lib.cdll()


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("fcvm_set_mesh", "fcvm_assemble", "fcvm_pcg_solve", "fcvm_update_stress_load",
                 "fcvm_update_peeq_csr", "fcvm_host_update_stress_load", "fcvm_host_solve", "fcvm_comm_init"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/fcvm_b200.h but not exported: {missing}"


def test_binding_table_covers_header():
    from fcvm_workbench_b200 import _lib
    assert sorted(_lib.exported_names()) == declared_symbols()


def test_no_cpu_fallback(lib):
    """Without a CUDA device the library refuses to create a context (and says so)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = ctypes.c_void_p()
    rc = lib.fcvm_create(ctypes.byref(ctx), 0)
    assert rc == -2
    assert b"no CPU fallback" in lib.fcvm_last_error()
    from fcvm_workbench_b200 import fcVM
    from fcvm_workbench_b200._lib import FcvmError
    from fcvm_workbench_b200.mesh import cube_model
    m = cube_model(1)
    with pytest.raises(FcvmError):
        fcVM.Engine(m.elNodes, m.nocoord, m.materialbyElement, m.fix)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "fcvm_workbench_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
