"""TEST INFRASTRUCTURE -- loads the UNMODIFIED reference module in this container.

``/root/reference/source code/fcVM.py`` imports FreeCAD, its FEM workbench,
pyvista, matplotlib and scikit-sparse at module level.  None of those exist
here, and none is needed by the numerical routines (``calcGSM``, ``calcTSM``,
``update_stress_load``, ``update_PEEQ_CSR``, ``mapStresses``, ``calcDisp``).  This
harness installs inert stand-ins for the GUI/CAD modules, a scipy-backed
stand-in for ``sksparse.cholmod.cholesky`` (same call protocol:
``factor = cholesky(A); x = factor(b)``, A given by its lower triangle as
CHOLMOD reads it) and then imports the reference file from where it lies.

It is used only (a) by ``oracle/gen_golden.py`` to write ``tests/golden/*.npz``
and (b) by the ``-m "not gpu"`` tests that cross-check ``oracle/fcvm_oracle.py``
when ``/root/reference`` is present.  Nothing on the product path imports it.

Stand-in for CHOLMOD: scikit-sparse 0.4.x (the reference's requirement) is not
installed and cannot be built here (needs SuiteSparse).  The direct solve is
replaced by SuperLU on the symmetrised matrix; both are backward-stable direct
solvers, results agree to round-off (cond * 1e-16).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from unittest import mock

# numba reads its configuration when first imported: point its on-disk cache away
# from the (read-only) reference tree before anything can import it.
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/fcvm_ref_numba_cache")

import numpy as np
import scipy.sparse as scsp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    """Where the unmodified reference lies: $FCVM_REFERENCE_ROOT, /root/reference (the build container), or
    ``oracle/_ref`` -- the git-ignored copy ``oracle/make_ref.py`` makes at build time so that the reference arm
    of ``bench.py`` can run on the GPU box, where /root/reference does not exist."""
    cands = [os.environ.get("FCVM_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "source code", "fcVM.py")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_root()
_SRC = os.path.join(REFERENCE_ROOT, "source code", "fcVM.py")


def available() -> bool:
    if not os.path.isfile(_SRC):
        return False
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return True


class _Factor:
    """``factor(b)`` protocol of sksparse.cholmod.Factor."""

    def __init__(self, A):
        A = scsp.csc_matrix(A)
        low = scsp.tril(A, format="csc")
        if (A - low).nnz == 0:          # lower triangle only: CHOLMOD's symmetric storage
            full = low + scsp.tril(A, k=-1, format="csc").T
        else:
            full = A
        self._lu = spla.splu(scsp.csc_matrix(full), permc_spec="MMD_AT_PLUS_A",
                             options=dict(SymmetricMode=True))

    def __call__(self, b):
        return self._lu.solve(np.asarray(b, dtype=np.float64))


def _cholesky(A, *a, **k):
    return _Factor(A)


_STUBS = ["FemGui", "FreeCAD", "FreeCADGui", "ObjectsFem", "Part", "pyvista",
          "matplotlib", "matplotlib.pyplot", "matplotlib.widgets", "matplotlib.ticker",
          "femtools", "femtools.membertools", "femmesh", "femmesh.meshsetsgetter",
          "femmesh.meshtools", "femresult", "femresult.resulttools", "feminout",
          "feminout.importToolsFem", "femtaskpanels", "femtaskpanels.task_result_mechanical"]

_module = None


def load():
    """Import the reference ``fcVM`` module (cached)."""
    global _module
    if _module is not None:
        return _module
    if not available():
        raise RuntimeError("reference sources (or numba) not present; golden fixtures are the only oracle pin here")
    for name in _STUBS:
        sys.modules.setdefault(name, mock.MagicMock(name=name))
    sk = types.ModuleType("sksparse")
    ch = types.ModuleType("sksparse.cholmod")
    ch.cholesky = _cholesky
    sk.cholmod = ch
    sys.modules.setdefault("sksparse", sk)
    sys.modules.setdefault("sksparse.cholmod", ch)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)          # dummyVM.py and fcVM.ini live at the root
    spec = importlib.util.spec_from_file_location("fcVM_reference", _SRC)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["fcVM_reference"] = mod              # numba's on-disk cache re-imports by name
    spec.loader.exec_module(mod)
    _module = mod
    return mod


def numba_fix(fix: dict):
    """The typed dictionary ``setUpInput`` builds (fcVM.py:222)."""
    from numba import types as nbt
    from numba.typed import Dict
    d = Dict.empty(key_type=nbt.int64, value_type=nbt.float64)
    for k, v in fix.items():
        d[int(k)] = float(v)
    return d


class _Window:
    """Inert stand-in for the Qt task panel (``fcVM_window``)."""

    def __init__(self, csr: bool):
        self.progressBar = mock.MagicMock()
        self.Step = mock.MagicMock()
        self.Load_Factor = mock.MagicMock()
        self.PEEQ = mock.MagicMock()
        self.CSR = mock.MagicMock()
        self.csrRbtn = mock.MagicMock()
        self.csrRbtn.isChecked.return_value = csr


def run_reference(model, ctl, clicks=()):
    """Run ``calcGSM`` + ``calcDisp`` of the reference on ``model`` / ``ctl``.

    ``clicks`` scripts the interactive load-displacement window: each entry is
    ``"stop"``, ``"add"`` , ``"rev"`` or ``("add", target_LF)``; when the list is exhausted the
    session stops (the "stop" button, fcVM.py:1659-1662).
    """
    ref = load()
    fix = numba_fix(model.fix)
    m = model
    out = ref.calcGSM(m.elNodes, m.nocoord, m.materialbyElement, fix, ctl.grav_x, ctl.grav_y, ctl.grav_z,
                      m.loadfaces, m.pressure, m.loadvertices, m.vertexloads, m.loadedges, m.edgeloads,
                      m.loadfaces_uni, m.faceloads)
    stm, row, col, glv, modf, V, lsx, lsy, lsz, ne, nn, x = out
    queue = list(clicks)

    def plot(fcVM, averaged, el_limit, ul_limit, un, lbd, csrplot, peeqmax, dl, du, target_LF, nstep, ue,
             *rest):
        if not queue:
            return False, dl, du, target_LF
        ev = queue.pop(0)
        tgt = target_LF
        if isinstance(ev, tuple):
            ev, tgt = ev
        if ev == "stop":
            return False, dl, du, tgt
        if ev == "rev":
            return True, -dl, -du, tgt
        if ev == "add":                                   # Index.add, fcVM.py:1664-1672
            LF = lbd[-1]
            if (target_LF - LF) * (tgt - LF) <= 0.0:
                dl = np.sign(tgt - LF) * 1.0 / nstep
                du = dl * ue
            return True, dl, du, tgt
        raise ValueError(ev)

    ref.plot = plot
    messages = []
    ref.prn_upd = lambda *a: messages.append("".join(str(o) for o in a))
    win = _Window(ctl.csr_option == "CSR")
    res = ref.calcDisp(m.elNodes, m.nocoord.copy(), m.fixdof, m.movdof, modf, m.materialbyElement, stm, row, col,
                       glv, ctl.nstep, ctl.iterat_max, ctl.error_max, ctl.relax, ctl.scale_re, ctl.scale_up,
                       ctl.scale_dn, ctl.sig_yield, ctl.disp_output, ctl.ultimate_strain, win, ctl.Et_E,
                       ctl.target_LF, x, m.noce, fix, ctl.grav_x, ctl.grav_y, ctl.grav_z, m.loadfaces, m.pressure,
                       m.loadvertices, m.vertexloads, m.loadedges, m.edgeloads, m.loadfaces_uni, m.faceloads,
                       ctl.gnl, ctl.maxImp, ctl.ev1, ctl.ev2)
    keys = ["displacements", "disp_el", "eigenval", "eigenvec", "stresses", "peeq", "sigmises", "csr", "lout",
            "un", "crip", "peeqplot", "pplot", "svmplot", "triaxplot", "ecrplot", "csrplot", "fail", "nocoord_old"]
    d = dict(zip(keys, res))
    d.update(stm=stm, row=row, col=col, glv=glv, modf=modf, V=V, loadsum=(lsx, lsy, lsz), x=x)
    d["messages"] = messages
    d["iters"] = iterations_per_step(messages)
    return d


def iterations_per_step(messages):
    """Newton iterations of every load step, read off the reference's own progress
    lines ("Step: n" / "Iteration: k, Error: e", fcVM.py:1315, 1344, 1455).  A restart
    resets the counter (fcVM.py:1484); the last value printed before the next "Step"
    is the count the step ended with."""
    its, cur = [], None
    for msg in messages:
        if msg.startswith("Step:") and "Load level" not in msg:
            if cur is not None:
                its.append(cur)
            cur = 0
        elif msg.startswith("Iteration:"):
            cur = int(msg.split(",")[0].split(":")[1])
    if cur is not None:
        its.append(cur)
    return its


class _StopTiming(Exception):
    pass


def time_reference(model, ctl, warmup: int, steps: int):
    """Newton iterations ``warmup+1 .. warmup+steps`` of the reference's own ``calcGSM`` + ``calcDisp`` on
    ``model`` / ``ctl``, timed from its progress lines ("Iteration: k, Error: e", fcVM.py:1455: one per Newton
    iteration).  Returns (seconds for the ``steps`` iterations, seconds of set-up before the first iteration:
    numba compilation, calcGSM, the sparse factorisation and the elastic solve)."""
    import time
    ref = load()
    fix = numba_fix(model.fix)
    m = model
    t0 = time.perf_counter()
    out = ref.calcGSM(m.elNodes, m.nocoord, m.materialbyElement, fix, ctl.grav_x, ctl.grav_y, ctl.grav_z,
                      m.loadfaces, m.pressure, m.loadvertices, m.vertexloads, m.loadedges, m.edgeloads,
                      m.loadfaces_uni, m.faceloads)
    stm, row, col, glv, modf, V, lsx, lsy, lsz, ne, nn, x = out
    stamps = []

    def prn(*a):
        msg = "".join(str(o) for o in a)
        if msg.startswith("Iteration:") and int(msg.split(",")[0].split(":")[1]) >= 1:
            stamps.append(time.perf_counter())
            if len(stamps) == warmup + steps:
                raise _StopTiming()

    ref.prn_upd = prn
    ref.plot = lambda *a, **k: (False,) + tuple(a[8:11])          # "stop" if the sweep ends first
    win = _Window(ctl.csr_option == "CSR")
    try:
        ref.calcDisp(m.elNodes, m.nocoord.copy(), m.fixdof, m.movdof, modf, m.materialbyElement, stm, row, col, glv,
                     ctl.nstep, ctl.iterat_max, ctl.error_max, ctl.relax, ctl.scale_re, ctl.scale_up, ctl.scale_dn,
                     ctl.sig_yield, ctl.disp_output, ctl.ultimate_strain, win, ctl.Et_E, ctl.target_LF, x, m.noce, fix,
                     ctl.grav_x, ctl.grav_y, ctl.grav_z, m.loadfaces, m.pressure, m.loadvertices, m.vertexloads,
                     m.loadedges, m.edgeloads, m.loadfaces_uni, m.faceloads, ctl.gnl, ctl.maxImp, ctl.ev1, ctl.ev2)
    except _StopTiming:
        pass
    if len(stamps) < warmup + steps:
        raise RuntimeError(f"the reference's load sweep ended after {len(stamps)} Newton iterations, fewer than "
                           f"warmup+steps = {warmup + steps}")
    return stamps[warmup + steps - 1] - stamps[warmup - 1], stamps[0] - t0
