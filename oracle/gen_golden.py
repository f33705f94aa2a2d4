"""TEST INFRASTRUCTURE -- writes tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference and numba):

    python oracle/gen_golden.py

Every fixture stores the inputs (the ``setUpInput`` arrays and the control
values) next to what the reference's own routines returned for them, so the
tests can replay the same inputs through the oracle and through the CUDA path
on machines where the reference does not exist.

Fixtures
--------
``tensile``        the reference's own model ``freeCAD files/tensile.FCStd`` with
                   ``control files/tensile.inp``; the scripted session reproduces the
                   rows of the committed ``output files/tensile.out``.
``cube2_platen``   48-element block, displacement control, hardening, one restart.
``cube2_force``    48-element block, traction + gravity, hardening, "add" click.
``cube2_gnly``     large-displacement branch (calcTSM every iteration).
``vm_uniaxial_tension`` the reference's VM_Uniaxial_Tension_Example (BASELINE config 0) with its control file.
``simple_shear``   the reference's Simple Shear model with its control file.
``embankment``     the reference's Embankment_with_Ditch_Example (BASELINE config 2) with its control file.
``cube2_maxrestarts`` force + gravity past the limit load with a small ``iterat_max``: the analysis ends with
                   MAXIMUM RESTARTS REACHED (fcVM.py:1460-1474: the last kept load level is overwritten).
``cube2_elastic``  nstep = 1: the linear-elastic analysis (no load stepping).
``column_buckling`` GNLY with imperfection: linear buckling (eigsh), imperfect geometry, restart.
``kernels``        single calls of calcGSM (element matrices), update_stress_load
                   (LD off/on), update_PEEQ_CSR and mapStresses on a distorted mesh
                   with random state.
"""
from __future__ import annotations

import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fcvm_workbench_b200.control import Control, read_control  # noqa: E402
from fcvm_workbench_b200.fcstd import read_fcstd  # noqa: E402
from fcvm_workbench_b200.mesh import cube_model  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def model_fields(m, prefix="m_"):
    fk = np.fromiter(m.fix.keys(), dtype=np.int64, count=len(m.fix))
    fv = np.fromiter(m.fix.values(), dtype=np.float64, count=len(m.fix))
    d = dict(name=np.array(m.name), elNodes=m.elNodes, nocoord=m.nocoord, fix_dof=fk, fix_val=fv, fixdof=m.fixdof,
             movdof=m.movdof, materialbyElement=m.materialbyElement[:1], noce=m.noce, loadfaces=m.loadfaces,
             pressure=m.pressure, loadvertices=m.loadvertices, vertexloads=m.vertexloads, loadedges=m.loadedges,
             edgeloads=m.edgeloads, loadfaces_uni=m.loadfaces_uni, faceloads=m.faceloads)
    return {prefix + k: v for k, v in d.items()}


def ctl_fields(c):
    return {"c_" + k: np.array(v) for k, v in dataclasses.asdict(c).items()}


def clicks_field(clicks):
    return np.array([f"{e[0]}:{e[1]}" if isinstance(e, tuple) else e for e in clicks] or [""], dtype="U32")


def analysis_case(name, m, c, clicks=(), extra=()):
    d = rh.run_reference(m, c, clicks=clicks)
    out = dict(model_fields(m))
    out.update(ctl_fields(c))
    out["clicks"] = clicks_field(clicks)
    for k in ("lout", "un", "crip", "peeqplot", "pplot", "svmplot", "triaxplot", "ecrplot", "csrplot",
              "displacements", "disp_el", "stresses", "peeq", "sigmises", "csr", "glv", "modf", "x"):
        out["r_" + k] = np.asarray(d[k])
    for k in extra:
        out["r_" + k] = np.asarray(d[k])
    out["r_iters"] = np.asarray(d["iters"], dtype=np.int64)
    out["r_V"] = np.array(d["V"])
    out["r_loadsum"] = np.array(d["loadsum"])
    # assembled lower-triangular CSC exactly as fcVM.py:1111 builds it
    import scipy.sparse as scsp
    ndof = len(d["glv"])
    gsm = scsp.csc_matrix((d["stm"], (d["row"], d["col"])), shape=(ndof, ndof))
    gsm.sum_duplicates()
    gsm.sort_indices()
    out["r_gsm_indptr"] = gsm.indptr.astype(np.int64)
    out["r_gsm_indices"] = gsm.indices.astype(np.int64)
    out["r_gsm_data"] = gsm.data
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: {m.ne} elements, {len(d['un'])} curve points, iters {d['iters']}")


def kernel_case():
    ref = rh.load()
    rng = np.random.default_rng(20240917)
    m = cube_model(2, size=4.0, mode="platen", top_disp=0.05, E=70000.0, nu=0.33, density=2.7e-6)
    nocoord = m.nocoord + rng.uniform(-0.06, 0.06, m.nocoord.shape)     # curved edges, distorted cells
    ne, nn = m.ne, m.nn
    fix = rh.numba_fix(m.fix)
    empty = rh.numba_fix({})
    L = (m.loadfaces, m.pressure, m.loadvertices, m.vertexloads, m.loadedges, m.edgeloads, m.loadfaces_uni,
         m.faceloads)
    out = dict(model_fields(m))
    out["m_nocoord"] = nocoord
    # element matrices: one-element models without constraints give esm's lower triangle in COO order
    esm = np.zeros((ne, 30, 30))
    for e in range(ne):
        loc = np.arange(1, 11, dtype=np.int64)[None, :]
        xyz = nocoord[m.elNodes[e] - 1]
        stm, row, col, *_ = ref.calcGSM(loc, xyz, m.materialbyElement[:1], empty, 0.0, 0.0, 0.0, *L)
        esm[e][row, col] = stm
        esm[e][col, row] = stm
    out["r_esm"] = esm
    # calcGSM with constraints and gravity
    g = (0.3, -0.2, -9.81)
    stm, row, col, glv, modf, V, lsx, lsy, lsz, _, _, x = ref.calcGSM(m.elNodes, nocoord, m.materialbyElement, fix,
                                                                      *g, *L)
    out.update(r_stm=stm, r_row=row, r_col=col, r_glv=glv, r_modf=modf, r_V=np.array(V), r_x=x, grav=np.array(g))
    gp10, _, _ = ref.gaussPoints()
    sy0 = 180.0
    for LD in (False, True):
        tag = "ld" if LD else "sm"
        du = rng.normal(0, 2.0e-3, 3 * nn)
        disp = rng.normal(0, 1.0e-2, 3 * nn)
        sig = rng.normal(0, 90.0, 24 * ne)
        sy = sy0 * (1.0 + 0.2 * rng.random(4 * ne))
        sig_new, sig_test = np.zeros(24 * ne), np.zeros(24 * ne)
        qin = np.zeros(3 * nn)
        pgp = np.full(4 * ne, False)
        Et_E = 0.02
        ref.update_stress_load(gp10, m.elNodes, nocoord, m.materialbyElement, sy, disp, du, sig, sig_new, sig_test,
                               qin, Et_E, LD, pgp)
        out.update({f"{tag}_du": du, f"{tag}_disp": disp, f"{tag}_sig": sig, f"{tag}_sy": sy,
                    f"{tag}_Et_E": np.array(Et_E), f"r_{tag}_sig_new": sig_new, f"r_{tag}_sig_test": sig_test,
                    f"r_{tag}_qin": qin, f"r_{tag}_pgp": pgp.copy()})
        if not LD:
            peeq = 1e-3 * rng.random(4 * ne)
            csr = 1e-2 * rng.random(4 * ne)
            triax, pres, svm, ecr = (np.zeros(4 * ne) for _ in range(4))
            out.update(pq_peeq0=peeq.copy(), pq_csr0=csr.copy(), pq_ult=np.array(0.25))
            sy2 = sy.copy()
            ref.update_PEEQ_CSR(ne, m.materialbyElement, sig_test, sig_new, sy2, 0.25, peeq, csr, triax, pres, svm,
                                ecr, Et_E)
            out.update(r_pq_sy=sy2, r_pq_peeq=peeq, r_pq_csr=csr, r_pq_triax=triax, r_pq_pressure=pres,
                       r_pq_sigmises=svm, r_pq_ecr=ecr)
            for averaged in (False, True):
                t = ref.mapStresses(averaged, m.elNodes, nocoord, sig_new, peeq, svm, csr, m.noce, sy0)
                k = "avg" if averaged else "max"
                out.update({f"r_map_{k}_stress": t[0], f"r_map_{k}_peeq": t[1], f"r_map_{k}_csr": t[2],
                            f"r_map_{k}_svm": t[3], f"r_map_{k}_triax": t[4]})
        else:
            # consistent tangent on the updated geometry with these plastic flags (calcTSM, nstep > 1)
            stm, _, _, row, col, glv2, modf2 = ref.calcTSM(8, m.elNodes, nocoord, m.materialbyElement, fix, *g, *L,
                                                           disp, du, sig, pgp, Et_E)
            out.update(r_tsm_stm=stm, r_tsm_row=row, r_tsm_col=col, r_tsm_glv=glv2, r_tsm_modf=modf2)
    np.savez_compressed(os.path.join(GOLD, "kernels.npz"), **out)
    print("kernels: written")


def main():
    os.makedirs(GOLD, exist_ok=True)
    if not rh.available():
        raise SystemExit("reference sources not present: fixtures can only be regenerated in the build container")
    m = read_fcstd(os.path.join(rh.REFERENCE_ROOT, "freeCAD files", "tensile.FCStd"))
    c = read_control(os.path.join(rh.REFERENCE_ROOT, "control files", "tensile.inp"))
    analysis_case("tensile", m, c, clicks=[("add", 0.4), ("add", 0.5), ("add", 0.6)])
    analysis_case("cube2_platen", cube_model(2, mode="platen", top_disp=0.1),
                  Control(sig_yield=240.0, nstep=8, error_max=1e-6, target_LF=2.0, Et_E=0.01))
    analysis_case("cube2_force", cube_model(2, mode="force", top_disp=300.0),
                  Control(sig_yield=240.0, nstep=8, error_max=1e-6, target_LF=1.5, Et_E=0.05, grav_z=-10.0),
                  clicks=["add"])
    analysis_case("cube2_gnly", cube_model(2, mode="platen", top_disp=0.4),
                  Control(sig_yield=240.0, nstep=6, error_max=1e-6, target_LF=2.0, Et_E=0.02, gnl="GNLY"))
    max_restarts_case()
    elastic_case()
    uniaxial_case()
    simple_shear_case()
    embankment_case()
    buckling_case()
    kernel_case()


def uniaxial_case():
    """BASELINE config 0: the reference's VM_Uniaxial_Tension_Example (FCStd + control file), continued past
    first yield with one "add" click (plateau at load factor 10 = yield stress / applied pressure)."""
    m = read_fcstd(os.path.join(rh.REFERENCE_ROOT, "freeCAD files", "VM_Uniaxial_Tension_Example.FCStd"))
    c = read_control(os.path.join(rh.REFERENCE_ROOT, "control files", "VM_Uniaxial_Tension_Example.inp"))
    analysis_case("vm_uniaxial_tension", m, c, clicks=[("add", 10.5)])


def simple_shear_case():
    """The reference's ``Simple Shear`` model with its control file (force-controlled shear, 96 Gauss points
    plastic at the end)."""
    m = read_fcstd(os.path.join(rh.REFERENCE_ROOT, "freeCAD files", "Simple Shear.FCStd"))
    c = read_control(os.path.join(rh.REFERENCE_ROOT, "control files", "Simple Shear.inp"))
    analysis_case("simple_shear", m, c)


def embankment_case():
    """BASELINE config 2: the reference's Embankment_with_Ditch_Example (659 elements, gravity-driven collapse
    of a soil body in plane strain) with its control file; reproduces the committed
    ``output files/Embankment_with_Ditch_Example.out``."""
    m = read_fcstd(os.path.join(rh.REFERENCE_ROOT, "freeCAD files", "Embankment_with_Ditch_Example.FCStd"))
    c = read_control(os.path.join(rh.REFERENCE_ROOT, "control files", "Embankment_with_Ditch_Example.inp"))
    analysis_case("embankment", m, c)


def buckling_case():
    """GNLY with an imperfection: linear buckling analysis (calcTSM nstep = 1, eigsh), imperfect geometry,
    restart and large-displacement load stepping (fcVM.py:1199-1294).  Pins the oracle's restatement of that
    branch; the CUDA path does not cover it yet."""
    m = cube_model(2, mode="platen", top_disp=-0.4, nxyz=(1, 1, 4), size=2.0)
    c = Control(sig_yield=240.0, nstep=4, error_max=1e-6, target_LF=1.0, Et_E=0.02, gnl="GNLY", maxImp="0.05",
                ev1="1.0", ev2="0.0")
    analysis_case("column_buckling", m, c, extra=("eigenval",))


def max_restarts_case():
    """The normal end of a collapse run: a step that does not converge within ``iterat_max`` after four
    restarts (fcVM.py:1457-1474).  The reference drops the failed load level, then overwrites the last kept
    one with ``lbd[step] + dl`` of the last Riks correction."""
    analysis_case("cube2_maxrestarts", cube_model(2, mode="force", top_disp=300.0),
                  Control(sig_yield=240.0, nstep=14, iterat_max=3, error_max=1e-8, target_LF=2.0, Et_E=0.0,
                          grav_z=-3.0e6))


def elastic_case():
    """nstep = 1: the reference's linear-elastic analysis (fcVM.py:1216-1223)."""
    analysis_case("cube2_elastic", cube_model(2, mode="force", top_disp=300.0),
                  Control(sig_yield=240.0, nstep=1, error_max=1e-6, target_LF=1.0, Et_E=0.0, grav_z=-10.0))


if __name__ == "__main__":
    main()
