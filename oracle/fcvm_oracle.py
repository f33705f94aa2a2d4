"""TEST INFRASTRUCTURE -- CPU oracle for the fcVM Newton / load-stepping path.

Python face of ``oracle/fcvm_oracle.c`` (the element routines, restated in plain
C from the reference's numba functions) plus a line-by-line restatement of the
load-stepping driver ``calcDisp`` (reference: source code/fcVM.py:1083-1635) for
the geometrically linear (GNLN) and large-displacement (GNLY, without the
eigen-buckling pre-analysis) branches.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
and ``--impl reference`` legs import this module.  The product
(``fcvm_workbench_b200``) never does.

Parity pin: ``oracle/gen_golden.py`` runs the UNMODIFIED reference through
``oracle/ref_harness.py``; its outputs are committed under ``tests/golden`` and
``tests/test_oracle_vs_golden.py`` holds this oracle to them (element matrices
and internal forces to 1e-12 relative, identical COO/CSC pattern, identical
plastic flags, load-displacement curves to 1e-9, equal iteration counts).

Direct solver: the reference factorises with CHOLMOD (scikit-sparse), which is
not installable here; SuperLU (scipy) on the symmetrised matrix stands in.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Callable, Optional

import numpy as np
import scipy.sparse as scsp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libfcvm_oracle.so")
_lib = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i16p = ctypes.POINTER(ctypes.c_int16)


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, no fast-math) into oracle/_build/."""
    src = os.path.join(_HERE, "fcvm_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.fcvm_oracle_calc_gsm.restype = ctypes.c_int64
        _lib.fcvm_oracle_calc_tsm.restype = ctypes.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _material(materialbyElement):
    m = np.asarray(materialbyElement, dtype=np.float64)
    return float(m[0][0]), float(m[0][1]), float(m[0][2])


def _fix_dense(fix, ndof):
    mask = np.zeros(ndof, dtype=np.uint8)
    val = np.zeros(ndof, dtype=np.float64)
    for d, v in fix.items():
        mask[d] = 1
        val[d] = v
    return mask, val


def load_vector(nocoord, loadfaces, pressure, loadvertices, vertexloads, loadedges, edgeloads, loadfaces_uni,
                faceloads, disp=None):
    """Surface / edge / vertex part of ``glv`` (fcVM.py:647-727, 856-938)."""
    nocoord = _c(nocoord, np.float64)
    nn = len(nocoord)
    glv = np.zeros(3 * nn)
    lf, pr = _c(loadfaces, np.int64), _c(pressure, np.float64)
    lv, vl = _c(loadvertices, np.int64), _c(vertexloads, np.float64)
    le, el_ = _c(loadedges, np.int64), _c(edgeloads, np.float64)
    lu, fl = _c(loadfaces_uni, np.int64), _c(faceloads, np.float64)
    d = _c(disp, np.float64) if disp is not None else None
    lib().fcvm_oracle_load_vector(ctypes.c_int64(nn), _p(nocoord, _f64p), _p(d, _f64p), ctypes.c_int64(len(pr)),
                                  _p(lf, _i64p), _p(pr, _f64p), ctypes.c_int64(len(lv)), _p(lv, _i64p),
                                  _p(vl, _f64p), ctypes.c_int64(len(le)), _p(le, _i64p), _p(el_, _f64p),
                                  ctypes.c_int64(len(lu)), _p(lu, _i64p), _p(fl, _f64p), _p(glv, _f64p))
    return glv


def calcGSM(elNodes, nocoord, materialbyElement, fix, grav_x, grav_y, grav_z, loadfaces, pressure, loadvertices,
            vertexloads, loadedges, edgeloads, loadfaces_uni, faceloads, return_esm=False):
    """fcVM.py:620-816.  Same return tuple as the reference (+ esm when asked)."""
    elNodes = _c(elNodes, np.int64)
    nocoord = _c(nocoord, np.float64)
    ne, nn = len(elNodes), len(nocoord)
    E, nu, rho = _material(materialbyElement)
    mask, val = _fix_dense(fix, 3 * nn)
    glv = load_vector(nocoord, loadfaces, pressure, loadvertices, vertexloads, loadedges, edgeloads,
                      loadfaces_uni, faceloads)
    ns = 465 * ne
    row = np.zeros(ns, dtype=np.int64)
    col = np.zeros(ns, dtype=np.int64)
    stm = np.zeros(ns)
    modf = np.zeros(3 * nn)
    x = np.zeros((4 * ne, 3))
    V = ctypes.c_double(0.0)
    esm = np.zeros((ne, 30, 30)) if return_esm else None
    pos = lib().fcvm_oracle_calc_gsm(
        ctypes.c_int64(ne), ctypes.c_int64(nn), _p(elNodes, _i64p), _p(nocoord, _f64p), ctypes.c_double(E),
        ctypes.c_double(nu), ctypes.c_double(rho), _p(mask, _u8p), _p(val, _f64p), ctypes.c_double(grav_x),
        ctypes.c_double(grav_y), ctypes.c_double(grav_z), _p(glv, _f64p), _p(row, _i64p), _p(col, _i64p),
        _p(stm, _f64p), _p(modf, _f64p), _p(x, _f64p), ctypes.byref(V), _p(esm, _f64p))
    row, col, stm = row[:pos], col[:pos], stm[:pos]
    ls = glv.reshape(-1, 3).sum(axis=0)
    out = (stm, row, col, glv, modf, V.value, ls[0], ls[1], ls[2], ne, nn, x)
    return out + (esm,) if return_esm else out


def calcTSM(nstep, elNodes, nocoord, materialbyElement, fix, grav_x, grav_y, grav_z, loadfaces, pressure,
            loadvertices, vertexloads, loadedges, edgeloads, loadfaces_uni, faceloads, disp_new, du, sig_old, pgp,
            Et_E, return_esm=False):
    """fcVM.py:819-1079: ``nstep > 1`` consistent tangent on the updated geometry; ``nstep == 1`` material and
    geometric stiffness of the linear-buckling analysis (returned as stms, stmg with full-matrix row/col)."""
    elNodes = _c(elNodes, np.int64)
    nocoord = _c(nocoord, np.float64)
    ne, nn = len(elNodes), len(nocoord)
    E, nu, rho = _material(materialbyElement)
    mask, val = _fix_dense(fix, 3 * nn)
    if not float(nstep) > 1.0:
        n9 = 900 * ne
        row, col = np.zeros(n9, dtype=np.int64), np.zeros(n9, dtype=np.int64)
        stms, stmg = np.zeros(n9), np.zeros(n9)
        lib().fcvm_oracle_calc_tsm_buckling.restype = ctypes.c_int64
        pos = lib().fcvm_oracle_calc_tsm_buckling(
            ctypes.c_int64(ne), _p(elNodes, _i64p), _p(nocoord, _f64p), ctypes.c_double(E), ctypes.c_double(nu),
            _p(mask, _u8p), _p(_c(sig_old, np.float64), _f64p), _p(_c(pgp, np.uint8), _u8p), ctypes.c_double(Et_E),
            _p(row, _i64p), _p(col, _i64p), _p(stms, _f64p), _p(stmg, _f64p))
        return None, stms[:pos], stmg[:pos], row[:pos], col[:pos], None, None
    disp_new = _c(disp_new, np.float64)
    glv = load_vector(nocoord, loadfaces, pressure, loadvertices, vertexloads, loadedges, edgeloads,
                      loadfaces_uni, faceloads, disp=disp_new)
    ns = 465 * ne
    row = np.zeros(ns, dtype=np.int64)
    col = np.zeros(ns, dtype=np.int64)
    stm = np.zeros(ns)
    modf = np.zeros(3 * nn)
    sig_old = _c(sig_old, np.float64)
    pg = _c(pgp, np.uint8)
    esm = np.zeros((ne, 30, 30)) if return_esm else None
    pos = lib().fcvm_oracle_calc_tsm(
        ctypes.c_int64(ne), ctypes.c_int64(nn), _p(elNodes, _i64p), _p(nocoord, _f64p), ctypes.c_double(E),
        ctypes.c_double(nu), ctypes.c_double(rho), _p(mask, _u8p), _p(val, _f64p), ctypes.c_double(grav_x),
        ctypes.c_double(grav_y), ctypes.c_double(grav_z), _p(disp_new, _f64p), _p(sig_old, _f64p), _p(pg, _u8p),
        ctypes.c_double(Et_E), _p(glv, _f64p), _p(row, _i64p), _p(col, _i64p), _p(stm, _f64p), _p(modf, _f64p),
        _p(esm, _f64p))
    out = (stm[:pos], None, None, row[:pos], col[:pos], glv, modf)
    return out + (esm,) if return_esm else out


def update_stress_load(gp10, elNodes, nocoord, materialbyElement, sig_yield, disp_new, du, sig, sig_update,
                       sig_test_global, qin, Et_E, LD, pgp):
    """fcVM.py:2196-2464.  Mutates sig_update, sig_test_global, qin (+=) and pgp like the reference."""
    elNodes = _c(elNodes, np.int64)
    nocoord = _c(nocoord, np.float64)
    E, nu, _ = _material(materialbyElement)
    ne, nn = len(elNodes), len(nocoord)
    pg = np.zeros(4 * ne, dtype=np.uint8)
    for a in (sig_update, sig_test_global, qin):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    lib().fcvm_oracle_update_stress_load(
        ctypes.c_int64(ne), ctypes.c_int64(nn), _p(elNodes, _i64p), _p(nocoord, _f64p), ctypes.c_double(E),
        ctypes.c_double(nu), _p(_c(sig_yield, np.float64), _f64p), _p(_c(disp_new, np.float64), _f64p),
        _p(_c(du, np.float64), _f64p), _p(_c(sig, np.float64), _f64p), _p(sig_update, _f64p),
        _p(sig_test_global, _f64p), _p(qin, _f64p), ctypes.c_double(Et_E), ctypes.c_int(1 if LD else 0),
        _p(pg, _u8p))
    pgp[:] = pg.astype(bool)


def update_PEEQ_CSR(nelem, materialbyElement, sig_test, sig_new, sig_yield, ultimate_strain, peeq, csr, triax,
                    pressure, sigmises, ecr, Et_E):
    """fcVM.py:2084-2137 (in-place on sig_yield, peeq, csr, triax, pressure, sigmises, ecr)."""
    E, nu, _ = _material(materialbyElement)
    for a in (sig_yield, peeq, csr, triax, pressure, sigmises, ecr):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    lib().fcvm_oracle_update_peeq_csr(
        ctypes.c_int64(nelem), ctypes.c_double(E), ctypes.c_double(nu), _p(_c(sig_test, np.float64), _f64p),
        _p(_c(sig_new, np.float64), _f64p), _p(sig_yield, _f64p), ctypes.c_double(ultimate_strain), _p(peeq, _f64p),
        _p(csr, _f64p), _p(triax, _f64p), _p(pressure, _f64p), _p(sigmises, _f64p), _p(ecr, _f64p),
        ctypes.c_double(Et_E))


def mapStresses(averaged, elNodes, nocoord, sig, peeq, sigvm, csr, noce, sig_yield):
    """fcVM.py:2496-2554."""
    elNodes = _c(elNodes, np.int64)
    ne, nn = len(elNodes), len(nocoord)
    t10s = np.zeros((nn, 6))
    t10p, t10c, t10v, t10t = np.zeros(nn), np.zeros(nn), np.zeros(nn), np.zeros(nn)
    lib().fcvm_oracle_map_stresses(
        ctypes.c_int(1 if averaged else 0), ctypes.c_int64(ne), ctypes.c_int64(nn), _p(elNodes, _i64p),
        _p(_c(sig, np.float64), _f64p), _p(_c(peeq, np.float64), _f64p), _p(_c(sigvm, np.float64), _f64p),
        _p(_c(csr, np.float64), _f64p), _p(_c(noce, np.int16), _i16p), ctypes.c_double(sig_yield), _p(t10s, _f64p),
        _p(t10p, _f64p), _p(t10c, _f64p), _p(t10v, _f64p), _p(t10t, _f64p))
    return t10s, t10p, t10c, t10v, t10t


class DirectFactor:
    """CHOLMOD stand-in: ``factor = DirectFactor(gsm); x = factor(b)`` (fcVM.py:1121, 1130)."""

    def __init__(self, lower_csc):
        low = scsp.csc_matrix(lower_csc)
        full = low + scsp.tril(low, k=-1, format="csc").T
        self._lu = spla.splu(scsp.csc_matrix(full), permc_spec="MMD_AT_PLUS_A", options=dict(SymmetricMode=True))

    def __call__(self, b):
        return self._lu.solve(np.asarray(b, dtype=np.float64))


def lower_csc(stm, row, col, ndof):
    """``scsp.csc_matrix((stm, (row, col)))`` of fcVM.py:1111 -- duplicates summed, pattern kept."""
    m = scsp.csc_matrix((stm, (row, col)), shape=(ndof, ndof))
    m.sum_duplicates()
    m.sort_indices()
    return m


class StopAnalysis(Exception):
    """Raised from an ``on_iteration`` hook to end the run early (bench.py's bounded sample)."""


def calcDisp(model, ctl, clicks=(), factorize: Optional[Callable] = None, log: Optional[Callable] = None,
             gsm_out=None, on_iteration: Optional[Callable] = None):
    """Load-stepping driver, restated from fcVM.py:1083-1635.

    ``clicks`` scripts the interactive window exactly as ``ref_harness.run_reference``.
    Returns a dict with the reference's return values plus per-step iteration
    counts (``iters``) and the plastic-flag history used by the parity tests.
    """
    m = model
    factorize = factorize or DirectFactor
    say = log or (lambda *a: None)
    fix = m.fix
    (stm, row, col, glv, modf, V, lsx, lsy, lsz, ne, nn, x) = calcGSM(
        m.elNodes, m.nocoord, m.materialbyElement, fix, ctl.grav_x, ctl.grav_y, ctl.grav_z, m.loadfaces,
        m.pressure, m.loadvertices, m.vertexloads, m.loadedges, m.edgeloads, m.loadfaces_uni, m.faceloads)
    elNodes, nocoord, mat = m.elNodes, m.nocoord.copy(), m.materialbyElement
    fixdof, movdof = m.fixdof, m.movdof
    nstep, iterat_max, error_max = ctl.nstep, ctl.iterat_max, ctl.error_max
    relax, scale_re, scale_up, scale_dn = ctl.relax, ctl.scale_re, ctl.scale_up, ctl.scale_dn
    disp_output, ultimate_strain, Et_E, target_LF = ctl.disp_output, ctl.ultimate_strain, ctl.Et_E, ctl.target_LF
    gnl, maxImp = ctl.gnl, float(ctl.maxImp)

    if gnl == "GNLY":                                       # fcVM.py:1087-1097
        LD = True
        relax = 1.0
        disp_output = "total"
        scale_up = 1.1
    else:
        LD = False
    eigenval, eigenvec = np.zeros(2), None

    ndof = len(glv)
    nelem = len(elNodes)
    nocoord_old = nocoord.copy()
    gsm = lower_csc(stm, row, col, ndof)                     # fcVM.py:1111
    if gsm_out is not None:
        gsm_out.append(gsm)
    qnorm = np.linalg.norm(glv)
    if qnorm < 1.0:
        qnorm = 1.0
    factor = factorize(gsm)                                  # fcVM.py:1121
    f = fixdof * glv + modf
    ue = factor(f)
    disp_el = ue.copy()

    dl0 = 1.0 / nstep
    dl = dl0
    du = dl * ue
    z24, z4 = (lambda: np.zeros(24 * nelem)), (lambda: np.zeros(4 * nelem))
    sig_new, sig_old, sig_test = z24(), z24(), z24()
    sig_yield = np.full(4 * nelem, ctl.sig_yield, dtype=np.float64)
    peeq, triax, pressure, sigmises, ecr, csr = z4(), z4(), z4(), z4(), z4(), z4()
    pgp = np.full(4 * nelem, False, dtype=bool)
    disp_new, disp_old = np.zeros(ndof), np.zeros(ndof)
    lbd = np.zeros(1)
    rfl = np.zeros(1)
    gp10 = None

    if max(movdof) == 1:                                     # fcVM.py:1169-1177
        qelastic = np.zeros(ndof)
        update_stress_load(gp10, elNodes, nocoord, mat, sig_yield, disp_new, ue, sig_old, sig_new, sig_test,
                           qelastic, Et_E, LD, pgp)
        qelastic *= movdof
        qnorm = np.linalg.norm(qelastic)
        sig_new = z24()

    step = -1
    cnt = True
    fail = False
    un, csrplot, crip, pplot, svmplot, triaxplot, peeqplot, peeqmax, ecrplot = (
        [0.], [0.], [0], [0.], [0.], [0.], [0.], [0.], [0.])
    lout = [0.]
    iters = []                                               # Newton iterations of every converged step
    nplastic = []

    update_stress_load(gp10, elNodes, nocoord, mat, 1.0e6 * sig_yield, np.zeros(ndof), ue, sig_old, sig_new,
                       sig_test, np.zeros(ndof), Et_E, False, pgp)          # fcVM.py:1195-1197

    if LD and not (float(nstep) > 1.0 and maxImp == 0.0):    # linear buckling analysis, fcVM.py:1199-1214
        from scipy.sparse.linalg import eigsh
        _, stms, stmg, brow, bcol, _, _ = calcTSM(1.0, elNodes, nocoord, mat, fix, ctl.grav_x, ctl.grav_y, ctl.grav_z,
                                                  m.loadfaces, m.pressure, m.loadvertices, m.vertexloads, m.loadedges,
                                                  m.edgeloads, m.loadfaces_uni, m.faceloads, np.zeros(ndof),
                                                  np.zeros(ndof), sig_new, pgp, Et_E)
        Kb = scsp.csc_matrix((stms, (brow, bcol)), shape=(ndof, ndof))
        Gb = -scsp.csc_matrix((stmg, (brow, bcol)), shape=(ndof, ndof))
        eigenval, eigenvec = eigsh(Kb, k=2, M=Gb, sigma=0.1, which="LM", mode="buckling")
        say(f"buckling load factors: {eigenval}")

    iterat_tot = 0
    mrr = False
    if float(nstep) != 1.0 and LD and maxImp != 0.0:         # imperfection and restart, fcVM.py:1224-1294
        ev1, ev2 = float(ctl.ev1), float(ctl.ev2)
        ua = ev1 / (ev1 + ev2) * eigenvec[:, 0] + ev2 / (ev1 + ev2) * eigenvec[:, 1]
        ub = ev1 / (ev1 + ev2) * eigenvec[:, 0] - ev2 / (ev1 + ev2) * eigenvec[:, 1]
        ma, mb = np.max(np.abs(ua)), np.max(np.abs(ub))
        if ma > mb:
            imax = np.argmax(np.abs(ua))
            imper = maxImp / ma * np.sign(ua[imax]) * ua
        else:
            imax = np.argmax(np.abs(ub))
            imper = maxImp / mb * np.sign(ub[imax]) * ub
        nocoord += imper.reshape(-1, 3)
        (stm, row, col, glv, modf, *_rest) = calcGSM(
            elNodes, nocoord, mat, fix, ctl.grav_x, ctl.grav_y, ctl.grav_z, m.loadfaces, m.pressure, m.loadvertices,
            m.vertexloads, m.loadedges, m.edgeloads, m.loadfaces_uni, m.faceloads)
        gsm = lower_csc(stm, row, col, ndof)
        qnorm = np.linalg.norm(glv)
        if qnorm < 1.0:
            qnorm = 1.0
        factor = factorize(gsm)
        f = fixdof * glv + modf
        ue = factor(f)
        disp_el = ue.copy()
        dl0 = 1.0 / nstep
        dl = dl0
        du = dl * ue
        sig_old, sig_test = z24(), z24()
        disp_new, disp_old = np.zeros(ndof), np.zeros(ndof)
        lbd = np.zeros(1)
    if float(nstep) == 1.0:                                  # elastic analysis, fcVM.py:1216-1223
        disp_new = ue
        lbd = np.append(lbd, 1.0)
        rfl = np.append(rfl, 1.0)
        un.append(float(np.max(np.abs(disp_new))))
        cnt = False
    sig_new = z24()                                          # fcVM.py:1300-1301: also wipes the elastic stresses of
    pgp = np.full(4 * nelem, False, dtype=bool)              # the call above, so nstep = 1 returns zero stresses
    queue = list(clicks)
    aa = 0.0

    def record():
        update_PEEQ_CSR(nelem, mat, sig_test, sig_new, sig_yield, ultimate_strain, peeq, csr, triax, pressure,
                        sigmises, ecr, Et_E)
        maxloc = int(np.argmax(csr))
        csrplot.append(np.max(csr))
        crip.append(maxloc)
        pplot.append(pressure[maxloc])
        svmplot.append(sigmises[maxloc])
        triaxplot.append(triax[maxloc])
        ecrplot.append(ecr[maxloc])
        peeqplot.append(peeq[maxloc])
        peeqmax.append(np.max(peeq))

    def un_now():
        d = disp_new[:3 * ((ndof - 1) // 3)].reshape(-1, 3)        # fcVM.py:1494-1497 (last node left out)
        return float(np.sqrt(np.max(np.sum(d * d, axis=1))))

    while cnt:
        cnt = False
        iRiks = True
        pstep = 0
        while pstep < nstep and not mrr:
            step += 1
            pstep += 1
            restart = 0
            say(f"Step: {step}")
            a = du.copy()
            if iRiks:
                sig_old = sig_new.copy()
                lbd = np.append(lbd, lbd[step] + dl)
            else:
                lbd[step + 1] = lbd[step] + dl
            qin = np.zeros(ndof)
            update_stress_load(gp10, elNodes, nocoord, mat, sig_yield, disp_new, du, sig_old, sig_new, sig_test,
                               qin, Et_E, LD, pgp)
            fex = fixdof * lbd[step + 1] * glv
            fin = fixdof * qin
            r = fex - fin
            rnorm = np.linalg.norm(r)
            error = rnorm / qnorm
            iterat = 0
            say(f"Iteration: {iterat}, Error: {error:.2e}")
            while error > error_max and not mrr:
                iterat += 1
                iterat_tot += 1
                if LD and (iterat == 1 or np.any(pgp)):                       # fcVM.py:1351-1396
                    stm, _, _, row, col, glv, modf = calcTSM(
                        nstep, elNodes, nocoord, mat, fix, ctl.grav_x, ctl.grav_y, ctl.grav_z, m.loadfaces,
                        m.pressure, m.loadvertices, m.vertexloads, m.loadedges, m.edgeloads, m.loadfaces_uni,
                        m.faceloads, disp_new, du, sig_old, pgp, Et_E)
                    tsm = lower_csc(stm, row, col, ndof)
                    try:
                        factor = factorize(tsm)
                    except Exception:
                        say("singular stiffness matrix")
                    fe = fixdof * glv + modf
                    ue = factor(fe)
                    a = ue.copy()
                    a *= np.linalg.norm(du) / np.linalg.norm(a)
                f = relax * r
                due = factor(f)
                if iRiks:                                                      # fcVM.py:1414-1421
                    dl = -np.dot(a, due) / np.dot(a, ue)
                    lbd[step + 1] += dl
                    aa = np.linalg.norm(a)
                else:
                    dl = 0.0
                du += due + dl * ue
                uu = np.linalg.norm(du)
                sf = min(aa / uu, 1.0)
                lbd[step + 1] = lbd[step] + sf * (lbd[step + 1] - lbd[step])
                du *= sf
                qin = np.zeros(ndof)
                update_stress_load(gp10, elNodes, nocoord, mat, sig_yield, disp_new, du, sig_old, sig_new,
                                   sig_test, qin, Et_E, LD, pgp)
                r = fixdof * (lbd[step + 1] * glv - qin)
                rnorm = np.linalg.norm(r)
                error = rnorm / qnorm
                say(f"Iteration: {iterat}, Error: {error:.2e}")
                if on_iteration is not None:
                    on_iteration(dict(step=step, iterat=iterat, iterat_tot=iterat_tot, error=error))
                if iterat > iterat_max:                                        # fcVM.py:1457-1484
                    say(f"RESTART # {restart + 1}")
                    if restart > 3:
                        say("MAXIMUM RESTARTS REACHED")
                        fail = False
                        step -= 1
                        lbd = lbd[:-1]
                        mrr = True
                    restart += 1
                    if step > 0 and not mrr:
                        dl = (lbd[step] - lbd[step - 1]) / scale_re / restart
                        du = (disp_new - disp_old) / scale_re / restart
                    elif not mrr:
                        dl = dl0 / scale_re / restart
                        du = dl * ue / scale_re / restart
                    lbd[step + 1] = lbd[step] + dl                             # unconditional, fcVM.py:1474
                    if not mrr:
                        qin = np.zeros(ndof)
                        update_stress_load(gp10, elNodes, nocoord, mat, sig_yield, disp_new, du, sig_old, sig_new,
                                           sig_test, qin, Et_E, LD, pgp)
                        r = fixdof * (lbd[step + 1] * (glv + modf) - qin)
                        rnorm = np.linalg.norm(r)
                        error = rnorm / qnorm
                        iterat = 0
            if abs(target_LF - lbd[step]) < abs(lbd[step + 1] - lbd[step]) and iRiks:   # fcVM.py:1486-1510
                say("REACHED TARGET LOAD")
                fac = (target_LF - lbd[step]) / (lbd[step + 1] - lbd[step])
                du = fac * du
                sig_new = sig_old + fac * (sig_new - sig_old)
                sig_test = sig_old + fac * (sig_test - sig_old)
                lbd[step + 1] = target_LF
                disp_new += du
                un.append(un_now())
                record()
                iters.append(iterat)
                nplastic.append(int(np.count_nonzero(pgp)))
                break
            elif not mrr:                                                               # fcVM.py:1515-1559
                disp_old = disp_new.copy()
                disp_new += du
                dl = lbd[step + 1] - lbd[step]
                if max(movdof) == 1:
                    rfl = np.append(rfl, np.sum(movdof * qin))
                if iterat > 10:
                    dl /= scale_dn
                    du /= scale_dn
                if iterat < 5:
                    dl *= scale_up
                    du *= scale_up
                un.append(un_now())
                record()
                iters.append(iterat)
                nplastic.append(int(np.count_nonzero(pgp)))
                if not iRiks:
                    break
        lout = rfl if max(movdof) == 1 else lbd
        # scripted stand-in for the interactive plot window (fcVM.py:1639-2080)
        if queue and not mrr:
            ev = queue.pop(0)
            tgt = target_LF
            if isinstance(ev, tuple):
                ev, tgt = ev
            if ev == "add":
                LF = lout[-1]
                if (target_LF - LF) * (tgt - LF) <= 0.0:
                    dl = np.sign(tgt - LF) * 1.0 / nstep
                    du = dl * ue
                cnt = True
            elif ev == "rev":
                dl, du, cnt = -dl, -du, True
            target_LF = tgt

    if disp_output == "total":
        dis = disp_new
    else:
        dis = disp_new - disp_old
    return dict(displacements=dis, disp_el=disp_el, stresses=sig_new, peeq=peeq, sigmises=sigmises, csr=csr,
                lout=np.asarray(lout), un=np.asarray(un), crip=np.asarray(crip), peeqplot=np.asarray(peeqplot),
                pplot=np.asarray(pplot), svmplot=np.asarray(svmplot), triaxplot=np.asarray(triaxplot),
                ecrplot=np.asarray(ecrplot), csrplot=np.asarray(csrplot), fail=fail, nocoord_old=nocoord_old,
                lbd=np.asarray(lbd), iters=np.asarray(iters), nplastic=np.asarray(nplastic), iterat_tot=iterat_tot,
                glv=glv, modf=modf, x=x, V=V, loadsum=(lsx, lsy, lsz), sig_yield=sig_yield, pgp=pgp,
                eigenval=eigenval, eigenvec=eigenvec,
                sig_test=sig_test, stm=stm, row=row, col=col)
