"""The reference-facing (HOST-buffer) face of the B200 path.

The reference keeps every array of the analysis in numpy on the host and calls
three heavy routines per Newton iteration (fcVM.py:1401 ``factor(f)``,
fcVM.py:1441 ``update_stress_load`` and, per load step, fcVM.py:1543
``update_PEEQ_CSR``).  A maintainer who binds ``libfcvm_b200.so`` inside fcVM.py
 gets exactly this data flow: host arrays in, GPU kernels, host
arrays out -- every call pays its PCIe copies.  ``HostEngine`` packages that flow
behind the method set ``fcVM.calcDisp`` drives, so the same load-stepping driver
runs either with device-resident state (``fcVM.Engine``) or through the
host-buffer C ABI (this class).  ``bench.py`` times the latter as its ``e2e`` leg.

Vectors are plain numpy arrays in page-locked memory (``fcvm_host_alloc``); the
cheap vector algebra between the heavy calls stays on the host, as in the reference
(the same arithmetic, run through torch's multi-threaded CPU kernels on views of the
same memory: at 4.1M dofs single-threaded numpy spent a third of the step there).
There is no CPU fallback for the heavy calls.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import call
from . import fcVM as _fc

_GP6 = (_fc.SIG_OLD, _fc.SIG_NEW, _fc.SIG_TEST)


def _timed(fn):
    """Accumulates the wall time spent in a method (``HostEngine.host_seconds``): where the host-buffer path spends
    its step -- PCIe copies + kernels inside the heavy calls, numpy algebra between them."""
    import functools
    import time

    @functools.wraps(fn)
    def wrap(self, *a, **k):
        t0 = time.perf_counter()
        try:
            return fn(self, *a, **k)
        finally:
            hs = self.__dict__.setdefault("host_seconds", {})
            hs[fn.__name__] = hs.get(fn.__name__, 0.0) + time.perf_counter() - t0
    return wrap


class HostEngine:
    """Same operator methods as ``fcVM.Engine``; handles are host numpy arrays."""

    def __init__(self, elNodes, nocoord, materialbyElement, fix=None, device: int = 0, comm=None):
        self.dev = _fc.Engine(elNodes, nocoord, materialbyElement, fix, device=device, comm=comm)
        self.ne, self.nn, self.ndof = self.dev.ne, self.dev.nn, self.dev.ndof
        self.comm = comm
        # element-partitioned run: host-side sums count shared dofs once and are completed over the ranks
        self._w = comm.part.interface(comm.rank)[0] if comm is not None and comm.world > 1 else None
        self._un_nodes = comm.part.un_nodes(comm.rank) if self._w is not None else None
        self._pinned = []
        self._tviews = {}
        self.host_seconds = {}
        self._bytes0 = self.dev.copy_bytes()
        ne = self.ne
        self._gp = {}
        for w in _GP6:
            self._gp[w] = self._alloc(24 * ne)
        for w in (_fc.SIG_YIELD, _fc.PEEQ, _fc.CSR, _fc.TRIAX, _fc.PRESSURE, _fc.SIGMISES, _fc.ECR):
            self._gp[w] = self._alloc(4 * ne)
        self._pgp = np.zeros(4 * ne, dtype=bool)
        self._nodal = {_fc.MODF: self._alloc(self.ndof), _fc.FIXDOF: self._alloc(self.ndof)}
        self._nodal[_fc.FIXDOF][:] = 1.0
        self._movdof = np.zeros(self.ndof)
        if fix is not None:
            self._set_masks()

    # -- memory ---------------------------------------------------------------------------------
    # PCIe traffic of this engine, counted inside the library where the copies are issued
    @property
    def h2d_bytes(self):
        return self.dev.copy_bytes()[0] - self._bytes0[0]

    @property
    def d2h_bytes(self):
        return self.dev.copy_bytes()[1] - self._bytes0[1]

    def _alloc(self, n: int) -> np.ndarray:
        p = ctypes.c_void_p()
        call("fcvm_host_alloc", 8 * int(n), ctypes.byref(p))
        self._pinned.append(p.value)
        a = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(int(n),))
        a[:] = 0.0
        return a

    def close(self):
        if self.dev is not None:
            self.dev.synchronize()
            self.dev.close()
            self.dev = None
            for p in self._pinned:
                _lib.cdll().fcvm_host_free(ctypes.c_void_p(p))
            self._pinned = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _set_masks(self):
        self._nodal[_fc.FIXDOF][:] = 1.0 - self.dev.fixmask
        self._movdof[:] = (self.dev.fixmask != 0) & (self.dev.fixval != 0.0)

    def set_deflation(self, *a, **k):
        return self.dev.set_deflation(*a, **k)

    def set_constraints(self, fix):
        self.dev.set_constraints(fix)
        self._set_masks()

    # -- vectors: numpy on the host, as in the reference ---------------------------------------------
    def vec(self, n: Optional[int] = None, host=None):
        v = self._alloc(self.ndof if n is None else n)
        if host is not None:
            v[:] = host
        return v

    def buf(self, which):
        return self._nodal[which]

    def put(self, v, host):
        v[:] = host

    def get(self, v, n=None):
        return np.array(v, copy=True)

    def zero(self, x, n=None):
        x[:] = 0.0

    @_timed
    def copy(self, x, y, n=None):
        y[:] = x

    # in-place forms without temporaries, on torch views of the same host memory (multi-threaded)
    def _tv(self, a):
        views = self.__dict__.setdefault("_tviews", {})
        v = views.get(a.ctypes.data)
        if v is None or v.numel() != a.size:
            v = views[a.ctypes.data] = torch.from_numpy(a)
        return v

    def _tmp(self, like):
        t = getattr(self, "_scratch", None)
        if t is None or t.shape != like.shape:
            t = self._scratch = torch.empty_like(like)
        return t

    @_timed
    def axpby(self, a, x, b, y, n=None):
        """y = a*x + b*y"""
        X, Y = self._tv(x), self._tv(y)
        if b == 0.0:
            torch.mul(X, a, out=Y)
        elif a == 0.0:
            Y.mul_(b)
        else:
            if b != 1.0:
                Y.mul_(b)
            Y.add_(X, alpha=a)

    @_timed
    def axpbypcz(self, a, x, b, y, c, z, n=None):
        """z = a*x + b*y + c*z"""
        X, Y, Z = self._tv(x), self._tv(y), self._tv(z)
        if c == 0.0:
            torch.mul(X, a, out=Z)
        else:
            if c != 1.0:
                Z.mul_(c)
            Z.add_(X, alpha=a)
        Z.add_(Y, alpha=b)

    def _gsum(self, v: float) -> float:
        return float(v) if self._w is None else float(sum(self.comm.allgather(float(v))))

    @_timed
    def dot(self, x, y, n=None):
        X, Y = self._tv(x), self._tv(y)
        if self._w is None:
            return float(torch.dot(X, Y))
        t = self._tmp(X)
        torch.mul(self._tv(self._w), X, out=t)
        return self._gsum(float(torch.dot(t, Y)))

    def norm(self, x):
        return float(np.sqrt(self.dot(x, x)))

    @_timed
    def residual(self, lbd, glv, qin, r):
        R = self._tv(r)
        torch.mul(self._tv(glv), lbd, out=R)
        R.sub_(self._tv(qin))
        R.mul_(self._tv(self._nodal[_fc.FIXDOF]))
        return self.norm(r)

    def masked_norm(self, x, mask_host):
        t = x * np.asarray(mask_host, dtype=np.float64)
        return float(np.sqrt(self._gsum(np.dot(t, t) if self._w is None else np.dot(self._w * t, t))))

    @_timed
    def max_node_disp(self, disp):
        nodes = (self.ndof - 1) // 3 if self._un_nodes is None else self._un_nodes      # fcVM.py:1494-1497
        d = disp[:3 * nodes].reshape(-1, 3)
        m = float(np.max(np.sum(d * d, axis=1))) if nodes else 0.0
        if self._w is not None:
            m = max(self.comm.allgather(m))
        return float(np.sqrt(m))

    @_timed
    def reaction(self, qin):
        return self._gsum(np.sum(self._movdof * qin) if self._w is None else np.sum(self._w * self._movdof * qin))

    # -- Gauss-point state on the host -------------------------------------------------------------------
    def gp_get(self, which):
        if which == _fc.PGP:
            return self._pgp.copy()
        return self._gp[which].copy()

    def gp_put(self, which, host):
        self._gp[which][:] = host

    def gp_fill(self, which, value):
        self._gp[which][:] = value

    @_timed
    def gp_copy(self, src, dst):
        self._gp[dst][:] = self._gp[src]

    def plastic_count(self):
        return int(round(self._gsum(np.count_nonzero(self._pgp))))

    @_timed
    def scale_step_stress(self, fac):
        so = self._gp[_fc.SIG_OLD]
        for w in (_fc.SIG_NEW, _fc.SIG_TEST):
            self._gp[w][:] = so + fac * (self._gp[w] - so)

    # -- the heavy calls: host buffers through the C ABI ---------------------------------------------------
    def assemble(self, glv=None, grav=(0.0, 0.0, 0.0), tangent=False, disp=None, Et_E: float = 0.0):
        """calcGSM / calcTSM + factorisation stand-in: the matrix stays on the device."""
        d = self.dev
        g = d.vec(host=glv) if glv is not None else None
        dd = d.vec(host=disp) if disp is not None else None
        if tangent:
            # SIG_OLD is uploaded here; the plastic flags are still resident from the last stress update
            d.gp_put(_fc.SIG_OLD, self._gp[_fc.SIG_OLD])
        d.assemble(g, grav, tangent=tangent, disp=dd, Et_E=Et_E)
        if glv is not None:
            glv[:] = d.get(g)
        self._nodal[_fc.MODF][:] = d.get(d.buf(_fc.MODF))
        for h in (g, dd):
            if h is not None:
                call("fcvm_vec_free", d._ctx, ctypes.c_void_p(h))
                d._vecs.remove(h)

    @_timed
    def solve(self, b, x, rtol=1e-10, max_iter=20000, use_x0=False, raise_on_noconv=True, recycle=False):
        """x = factor(b) (fcVM.py:1130, 1401): one h2d of b, PCG on the device, one d2h of x."""
        self.dev.host_solve(b, rtol, max_iter, out=x, raise_on_noconv=raise_on_noconv, recycle=recycle)
        self.last_solve = self.dev.last_solve
        return self.last_solve

    @_timed
    def update_stress_load(self, disp_new, du, qin, Et_E, LD=False, yield_scale=1.0):
        """update_stress_load(...) of fcVM.py:2196 with the reference's host arrays."""
        g = self._gp
        sy = g[_fc.SIG_YIELD] if yield_scale == 1.0 else yield_scale * g[_fc.SIG_YIELD]
        qin[:] = 0.0                                                     # the reference passes zeros (fcVM.py:1324)
        self.dev.host_update_stress_load(sy, disp_new, du, g[_fc.SIG_OLD], g[_fc.SIG_NEW], g[_fc.SIG_TEST], qin,
                                         Et_E, LD, self._pgp)

    @_timed
    def update_peeq_csr(self, ultimate_strain, Et_E):
        g = self._gp
        res = _fc.update_PEEQ_CSR(self.ne, None, g[_fc.SIG_TEST], g[_fc.SIG_NEW], g[_fc.SIG_YIELD], ultimate_strain,
                                  g[_fc.PEEQ], g[_fc.CSR], g[_fc.TRIAX], g[_fc.PRESSURE], g[_fc.SIGMISES],
                                  g[_fc.ECR], Et_E, engine=self.dev)
        return res

    def map_stresses(self, averaged, sig_yield, noce=None):
        d, g = self.dev, self._gp
        for w in (_fc.SIG_NEW, _fc.PEEQ, _fc.SIGMISES, _fc.CSR):
            d.gp_put(w, g[w])
        return d.map_stresses(averaged, sig_yield, noce)

    # -- measurement -----------------------------------------------------------------------------------------
    def synchronize(self):
        self.dev.synchronize()

    def launch_count(self):
        return self.dev.launch_count()
