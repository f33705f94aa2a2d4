"""ctypes binding of ``libfcvm_b200.so`` (the C ABI declared in include/fcvm_b200.h).

There is no CPU fallback: if the library is missing it must be built
(``python -m fcvm_workbench_b200.build``); if no CUDA device is usable every
compute call raises ``FcvmError``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int16, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FCVM_LIB_PATH") or os.path.join(_HERE, "libfcvm_b200.so")   # override: kernel experiments

f64p = POINTER(c_double)
i64p = POINTER(c_int64)
u8p = POINTER(c_uint8)
i16p = POINTER(c_int16)
ctxp = c_void_p


class FcvmError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"fcvm_b200 error {code}: {message}")
        self.code = code


E_NOCONV = -4
E_INDEFINITE = -6

# name -> (argtypes); every function returns int unless listed in _RESTYPE
_SIGNATURES = {
    "fcvm_create": [POINTER(ctxp), c_int],
    "fcvm_destroy": [ctxp],
    "fcvm_set_stream": [ctxp, c_void_p],
    "fcvm_synchronize": [ctxp],
    "fcvm_set_mesh": [ctxp, c_int64, c_int64, i64p, f64p, c_double, c_double, c_double],
    "fcvm_set_constraints": [ctxp, u8p, f64p],
    "fcvm_set_coordinates": [ctxp, f64p],
    "fcvm_assemble_buckling": [ctxp, c_double],
    "fcvm_spmv_geometric": [ctxp, c_void_p, c_void_p],
    "fcvm_set_interface": [ctxp, f64p, c_int64, i64p, i64p, c_int64],
    "fcvm_vec_alloc": [ctxp, c_int64, POINTER(c_void_p)],
    "fcvm_vec_free": [ctxp, c_void_p],
    "fcvm_buf": [ctxp, c_int, POINTER(c_void_p), i64p],
    "fcvm_h2d": [ctxp, c_void_p, c_void_p, c_int64],
    "fcvm_d2h": [ctxp, c_void_p, c_void_p, c_int64],
    "fcvm_vec_zero": [ctxp, c_int64, c_void_p],
    "fcvm_vec_copy": [ctxp, c_int64, c_void_p, c_void_p],
    "fcvm_vec_axpby": [ctxp, c_int64, c_double, c_void_p, c_double, c_void_p],
    "fcvm_vec_axpbypcz": [ctxp, c_int64, c_double, c_void_p, c_double, c_void_p, c_double, c_void_p],
    "fcvm_vec_dot": [ctxp, c_int64, c_void_p, c_void_p, f64p],
    "fcvm_residual": [ctxp, c_double, c_void_p, c_void_p, c_void_p, f64p],
    "fcvm_max_node_disp": [ctxp, c_void_p, f64p],
    "fcvm_reaction": [ctxp, c_void_p, f64p],
    "fcvm_gp_to_host": [ctxp, c_void_p, c_int, f64p],
    "fcvm_gp_from_host": [ctxp, f64p, c_int, c_void_p],
    "fcvm_gp_fill": [ctxp, c_int, c_double],
    "fcvm_pgp_to_host": [ctxp, u8p],
    "fcvm_pgp_count": [ctxp, i64p],
    "fcvm_assemble": [ctxp, c_int, c_void_p, c_double, c_double, c_double, c_double, c_void_p],
    "fcvm_element_matrices": [ctxp, c_int, c_void_p, c_double, c_void_p],
    "fcvm_export_csc_lower": [ctxp, i64p, i64p, i64p, f64p],
    "fcvm_spmv": [ctxp, c_void_p, c_void_p],
    "fcvm_matfree_apply": [ctxp, c_void_p, c_void_p],
    "fcvm_set_deflation": [ctxp, c_int, c_int, c_int, POINTER(ctypes.c_int32), f64p, f64p, u8p],
    "fcvm_deflation_stats": [ctxp, i64p, i64p],
    "fcvm_pcg_solve": [ctxp, c_void_p, c_void_p, c_double, c_int, c_int, POINTER(c_int), f64p],
    "fcvm_update_stress_load": [ctxp, c_void_p, c_void_p, c_void_p, c_double, c_int, c_double],
    "fcvm_update_peeq_csr": [ctxp, c_double, c_double, i64p, f64p],
    "fcvm_scale_step_stress": [ctxp, c_double],
    "fcvm_map_stresses": [ctxp, c_int, c_double, i16p, f64p, f64p, f64p, f64p, f64p],
    "fcvm_comm_unique_id": [c_void_p],
    "fcvm_comm_init": [ctxp, c_void_p, c_int, c_int],
    "fcvm_comm_allreduce_sum": [ctxp, c_void_p, c_int64],
    "fcvm_comm_allreduce_max": [ctxp, c_void_p, c_int64],
    "fcvm_set_un_nodes": [ctxp, c_int64],
    "fcvm_p2p_create": [ctxp, c_int64, c_int64, c_void_p],
    "fcvm_p2p_attach": [ctxp, c_void_p, c_int, POINTER(ctypes.c_int32), POINTER(ctypes.c_int32), POINTER(ctypes.c_int32),
                        i64p, c_int, POINTER(ctypes.c_int32), POINTER(ctypes.c_int32), i64p],
    "fcvm_p2p_interface_sum": [ctxp, c_void_p],
    "fcvm_interface_sum": [ctxp, c_void_p],
    "fcvm_host_alloc": [c_int64, POINTER(c_void_p)],
    "fcvm_host_free": [c_void_p],
    "fcvm_host_update_stress_load": [ctxp, f64p, f64p, f64p, f64p, f64p, f64p, f64p, c_double, c_int, u8p],
    "fcvm_host_solve": [ctxp, f64p, f64p, c_double, c_int, c_int, POINTER(c_int), f64p],
    "fcvm_timer_start": [ctxp],
    "fcvm_timer_stop_ms": [ctxp, POINTER(c_float)],
    "fcvm_profile_enable": [ctxp, c_int],
    "fcvm_profile_get": [ctxp, c_int, f64p, i64p],
    "fcvm_profile_reset": [ctxp],
    "fcvm_profile_seen": [ctxp, c_int],
    "fcvm_copy_bytes": [ctxp, i64p, i64p],
    "fcvm_matrix_stats": [ctxp, i64p, i64p, i64p],
    "fcvm_last_error": [],
    "fcvm_version": [],
    "fcvm_num_elements": [ctxp],
    "fcvm_num_nodes": [ctxp],
    "fcvm_launch_count": [ctxp],
}
_RESTYPE = {"fcvm_last_error": c_char_p, "fcvm_num_elements": c_int64, "fcvm_num_nodes": c_int64,
            "fcvm_launch_count": c_int64, "fcvm_profile_seen": c_int64}
_NO_CHECK = {"fcvm_last_error", "fcvm_version", "fcvm_num_elements", "fcvm_num_nodes", "fcvm_launch_count",
             "fcvm_profile_seen"}

_cdll = None


def exported_names():
    return sorted(_SIGNATURES)


def cdll():
    """The raw library (loaded once).  Raises if it has not been built."""
    global _cdll
    if _cdll is None:
        if not os.path.isfile(LIB_PATH):
            raise FcvmError(-2, f"{LIB_PATH} not built: run `python -m fcvm_workbench_b200.build` "
                                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, args in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, c_int)
        _cdll = lib
    return _cdll


def call(name, *args, allow=()):
    """Call an entry point and raise FcvmError on a non-zero return code."""
    lib = cdll()
    rc = getattr(lib, name)(*args)
    if name in _NO_CHECK:
        return rc
    if rc != 0 and rc not in allow:
        raise FcvmError(rc, lib.fcvm_last_error().decode(errors="replace"))
    return rc
