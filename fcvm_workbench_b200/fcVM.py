"""Host side of the B200 path: the reference's routines, same names, same arguments.

This module mirrors the operator interface of the reference's ``source code/fcVM.py`` for
the Newton-Raphson load-stepping path:

==========================  =====================  ==========================================
reference                   fcVM.py lines          here
==========================  =====================  ==========================================
``calcGSM``                 620-816                ``calcGSM`` / ``Engine.assemble``
``calcTSM`` (nstep > 1)     819-1079               ``Engine.assemble(tangent=True)``
``cholesky`` + ``factor``   1121-1135, 1401        ``Engine.solve`` (PCG on the device)
``update_stress_load``      2196-2464              ``update_stress_load`` / ``Engine.update_stress_load``
``update_PEEQ_CSR``         2084-2137              ``Engine.update_peeq_csr``
``mapStresses``             2496-2554              ``mapStresses``
``calcDisp``                1083-1635              ``calcDisp``
==========================  =====================  ==========================================

All arithmetic runs in ``libfcvm_b200.so`` (hand-written sm_100a kernels) through the C ABI of
``include/fcvm_b200.h``; state stays resident on the GPU between calls.  There is no CPU
fallback: without the library or a CUDA device the calls raise ``FcvmError``.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

from . import _lib
from ._lib import FcvmError, call, f64p, i16p, i64p, u8p
from .loads import surface_load_vector

# named device buffers (include/fcvm_b200.h)
SIG_OLD, SIG_NEW, SIG_TEST, SIG_YIELD, PEEQ, CSR, TRIAX, PRESSURE, SIGMISES, ECR, PGP, MODF, GLV, FIXDOF = range(14)


AUTO_DEFLATION = 6144      # coarse unknowns calcDisp asks for by default on large meshes


class StopAnalysis(Exception):
    """Raised from an ``on_iteration`` hook to end the analysis early (a scripted "stop" click)."""


def _np(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _ptr(a, t):
    return a.ctypes.data_as(t)


class Engine:
    """A mesh resident on one GPU: connectivity, sparsity pattern, Gauss-point state."""

    def __init__(self, elNodes, nocoord, materialbyElement, fix=None, device: int = 0, stream: Optional[int] = None,
                 comm=None):
        self._ctx = ctypes.c_void_p()
        call("fcvm_create", ctypes.byref(self._ctx), int(device))
        if stream is not None:
            call("fcvm_set_stream", self._ctx, ctypes.c_void_p(stream))
        el = _np(elNodes, np.int64)
        xyz = _np(nocoord, np.float64)
        mat = np.asarray(materialbyElement, dtype=np.float64)
        self.ne, self.nn = int(el.shape[0]), int(xyz.shape[0])
        self.ndof = 3 * self.nn
        self._elNodes, self._nocoord = el, xyz
        self.deflation_grid = None
        self.E, self.nu, self.density = float(mat[0][0]), float(mat[0][1]), float(mat[0][2])
        call("fcvm_set_mesh", self._ctx, self.ne, self.nn, _ptr(el, i64p), _ptr(xyz, f64p), self.E, self.nu,
             self.density)
        self._vecs = []
        self.comm = comm
        if comm is not None:
            comm.attach(self)
        if fix is not None:
            self.set_constraints(fix)
        if os.environ.get("FCVM_DEFLATION"):                      # e.g. FCVM_DEFLATION=6144 (coarse unknowns)
            self.set_deflation(int(os.environ["FCVM_DEFLATION"]))

    # -- lifetime ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            for v in self._vecs:
                _lib.cdll().fcvm_vec_free(self._ctx, ctypes.c_void_p(v))
            self._vecs = []
            _lib.cdll().fcvm_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- setup ----------------------------------------------------------------------------
    def set_constraints(self, fix):
        """``fix``: the reference's dof -> value dictionary (fcVM.py:222) or (mask, value) arrays."""
        if isinstance(fix, tuple):
            mask, val = _np(fix[0], np.uint8), _np(fix[1], np.float64)
        else:
            mask = np.zeros(self.ndof, dtype=np.uint8)
            val = np.zeros(self.ndof, dtype=np.float64)
            for d, v in fix.items():
                mask[int(d)] = 1
                val[int(d)] = float(v)
        self.fixmask, self.fixval = mask, val
        call("fcvm_set_constraints", self._ctx, _ptr(mask, u8p), _ptr(val, f64p))

    def set_deflation(self, target_unknowns: int = 3072, grid=None):
        """Switch on the second preconditioner level: rigid-body-mode deflation over box clusters of
        nodes (csrc/fcvm_deflation.cu).  ``target_unknowns`` bounds the dense coarse problem (six per
        cluster); ``grid=(ncx, ncy, ncz)`` overrides the automatic choice; ``target_unknowns=0``
        switches it off.  Takes effect at the next ``assemble``.  Returns the grid used."""
        if not target_unknowns and grid is None:
            call("fcvm_set_deflation", self._ctx, 0, 0, 0, None, None, None, None)
            self.deflation_grid = None
            return None
        xyz, el = self._nocoord, self._elNodes
        lo, hi = xyz.min(axis=0), xyz.max(axis=0)
        ext = np.ptp(xyz[el - 1], axis=1).max(axis=0)                 # widest element per direction
        if self.comm is not None and self.comm.world > 1:
            every = self.comm.allgather((lo, hi, ext))
            lo = np.min([e[0] for e in every], axis=0)
            hi = np.max([e[1] for e in every], axis=0)
            ext = np.max([e[2] for e in every], axis=0)
        grid, h = deflation_boxes(lo, hi, ext, target_unknowns, grid)
        ijk = np.minimum(((xyz - lo) / h).astype(np.int64), grid - 1)
        cid = np.ascontiguousarray(ijk[:, 0] + grid[0] * (ijk[:, 1] + grid[1] * ijk[:, 2]), dtype=np.int32)
        # a box needs a handful of nodes with all three dofs free for its six modes to be independent
        # (otherwise the dense coarse matrix is singular); boxes below that carry no modes
        free_node = (self.fixmask.reshape(-1, 3) == 0).all(axis=1) if getattr(self, "fixmask", None) is not None \
            else np.ones(self.nn, dtype=bool)
        if self.comm is not None and self.comm.world > 1:
            free_node = free_node & (self.comm.part.multiplicity[self.comm.part.nodes[self.comm.rank]] == 1)
        count = np.bincount(cid[free_node], minlength=int(np.prod(grid))).astype(np.int64)
        if self.comm is not None and self.comm.world > 1:
            count = np.sum(self.comm.allgather(count), axis=0)
        active = np.ascontiguousarray(count >= 8, dtype=np.uint8)
        lo_c, h_c = _np(lo, np.float64), _np(h, np.float64)
        call("fcvm_set_deflation", self._ctx, int(grid[0]), int(grid[1]), int(grid[2]),
             cid.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _ptr(lo_c, f64p), _ptr(h_c, f64p), _ptr(active, u8p))
        self.deflation_active_boxes = int(active.sum())
        self.deflation_grid = tuple(int(g) for g in grid)
        return self.deflation_grid

    # -- device vectors ---------------------------------------------------------------------
    def vec(self, n: Optional[int] = None, host=None) -> int:
        """Allocate a zeroed device vector (default length 3*nn); returns the device address."""
        n = self.ndof if n is None else int(n)
        p = ctypes.c_void_p()
        call("fcvm_vec_alloc", self._ctx, n, ctypes.byref(p))
        self._vecs.append(p.value)
        if host is not None:
            self.put(p.value, host)
        return p.value

    def buf(self, which: int) -> int:
        p = ctypes.c_void_p()
        n = ctypes.c_int64()
        call("fcvm_buf", self._ctx, which, ctypes.byref(p), ctypes.byref(n))
        return p.value

    def put(self, dev: int, host):
        h = _np(host, np.float64)
        call("fcvm_h2d", self._ctx, ctypes.c_void_p(dev), h.ctypes.data_as(ctypes.c_void_p), h.nbytes)

    def get(self, dev: int, n: Optional[int] = None) -> np.ndarray:
        out = np.empty(self.ndof if n is None else int(n), dtype=np.float64)
        call("fcvm_d2h", self._ctx, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(dev), out.nbytes)
        return out

    def zero(self, x, n=None):
        call("fcvm_vec_zero", self._ctx, self.ndof if n is None else n, ctypes.c_void_p(x))

    def copy(self, x, y, n=None):
        call("fcvm_vec_copy", self._ctx, self.ndof if n is None else n, ctypes.c_void_p(x), ctypes.c_void_p(y))

    def axpby(self, a, x, b, y, n=None):
        """y = a*x + b*y"""
        call("fcvm_vec_axpby", self._ctx, self.ndof if n is None else n, float(a), ctypes.c_void_p(x), float(b),
             ctypes.c_void_p(y))

    def axpbypcz(self, a, x, b, y, c, z, n=None):
        """z = a*x + b*y + c*z"""
        call("fcvm_vec_axpbypcz", self._ctx, self.ndof if n is None else n, float(a), ctypes.c_void_p(x), float(b),
             ctypes.c_void_p(y), float(c), ctypes.c_void_p(z))

    def dot(self, x, y, n=None) -> float:
        out = ctypes.c_double()
        call("fcvm_vec_dot", self._ctx, self.ndof if n is None else n, ctypes.c_void_p(x), ctypes.c_void_p(y),
             ctypes.byref(out))
        return out.value

    def norm(self, x) -> float:
        return float(np.sqrt(self.dot(x, x)))

    def residual(self, lbd, glv, qin, r) -> float:
        """r = fixdof*(lbd*glv - qin); returns ||r|| (fcVM.py:1329-1338, 1446-1447)."""
        out = ctypes.c_double()
        call("fcvm_residual", self._ctx, float(lbd), ctypes.c_void_p(glv), ctypes.c_void_p(qin), ctypes.c_void_p(r),
             ctypes.byref(out))
        return out.value

    def masked_norm(self, x, mask_host) -> float:
        """||mask * x||_2 for a host 0/1 mask (used once per analysis, fcVM.py:1174-1176)."""
        t = self.vec(host=self.get(x) * np.asarray(mask_host, dtype=np.float64))
        out = self.norm(t)
        call("fcvm_vec_free", self._ctx, ctypes.c_void_p(t))
        self._vecs.remove(t)
        return out

    def max_node_disp(self, disp) -> float:
        out = ctypes.c_double()
        call("fcvm_max_node_disp", self._ctx, ctypes.c_void_p(disp), ctypes.byref(out))
        return out.value

    def reaction(self, qin) -> float:
        out = ctypes.c_double()
        call("fcvm_reaction", self._ctx, ctypes.c_void_p(qin), ctypes.byref(out))
        return out.value

    # -- Gauss-point state --------------------------------------------------------------------
    def gp_get(self, which: int) -> np.ndarray:
        """Gauss-point array in the reference layout (24*ne stresses or 4*ne scalars)."""
        if which == PGP:
            out = np.empty(4 * self.ne, dtype=np.uint8)
            call("fcvm_pgp_to_host", self._ctx, _ptr(out, u8p))
            return out.astype(bool)
        ncomp = 6 if which <= SIG_TEST else 1
        out = np.empty(4 * self.ne * ncomp, dtype=np.float64)
        call("fcvm_gp_to_host", self._ctx, ctypes.c_void_p(self.buf(which)), ncomp, _ptr(out, f64p))
        return out

    def gp_put(self, which: int, host):
        ncomp = 6 if which <= SIG_TEST else 1
        h = _np(host, np.float64)
        assert h.size == 4 * self.ne * ncomp
        call("fcvm_gp_from_host", self._ctx, _ptr(h, f64p), ncomp, ctypes.c_void_p(self.buf(which)))

    def gp_fill(self, which: int, value: float):
        call("fcvm_gp_fill", self._ctx, which, float(value))

    def gp_copy(self, src: int, dst: int):
        n = (24 if src <= SIG_TEST else 4) * self.ne
        self.copy(self.buf(src), self.buf(dst), n)

    def plastic_count(self) -> int:
        out = ctypes.c_int64()
        call("fcvm_pgp_count", self._ctx, ctypes.byref(out))
        return out.value

    # -- operators ------------------------------------------------------------------------------
    def assemble(self, glv: Optional[int] = None, grav=(0.0, 0.0, 0.0), tangent=False, disp: Optional[int] = None,
                 Et_E: float = 0.0):
        """Element integration + deterministic assembly + constraint elimination.

        ``glv`` (device, holding the surface loads) receives the gravity load; ``modf`` is left
        in the named buffer MODF.
        """
        try:
            call("fcvm_assemble", self._ctx, 1 if tangent else 0, ctypes.c_void_p(disp) if disp else None, float(Et_E),
                 float(grav[0]), float(grav[1]), float(grav[2]), ctypes.c_void_p(glv) if glv else None)
        except FcvmError as e:
            if "deflation" not in str(e):
                raise
            # the matrix itself is assembled; only the coarse level could not be built (e.g. a box with
            # fewer free dofs than rigid-body modes): carry on with block-Jacobi PCG alone
            import warnings
            warnings.warn(f"deflation switched off: {e}")
            self.set_deflation(0)
        if glv and self.comm is not None and self.comm.world > 1:
            self.interface_sum(glv)

    def element_matrices(self, tangent=False, disp: Optional[int] = None, Et_E: float = 0.0) -> np.ndarray:
        d = self.vec(900 * self.ne)
        call("fcvm_element_matrices", self._ctx, 1 if tangent else 0, ctypes.c_void_p(disp) if disp else None,
             float(Et_E), ctypes.c_void_p(d))
        out = self.get(d, 900 * self.ne).reshape(self.ne, 30, 30)
        call("fcvm_vec_free", self._ctx, ctypes.c_void_p(d))
        self._vecs.remove(d)
        return out

    def export_csc_lower(self):
        """(indptr, indices, data) of the lower-triangular CSC matrix scipy builds at fcVM.py:1111."""
        nnz = ctypes.c_int64()
        call("fcvm_export_csc_lower", self._ctx, ctypes.byref(nnz), None, None, None)
        indptr = np.empty(self.ndof + 1, dtype=np.int64)
        indices = np.empty(nnz.value, dtype=np.int64)
        data = np.empty(nnz.value, dtype=np.float64)
        call("fcvm_export_csc_lower", self._ctx, ctypes.byref(nnz), _ptr(indptr, i64p), _ptr(indices, i64p),
             _ptr(data, f64p))
        return indptr, indices, data

    def spmv(self, x, y):
        call("fcvm_spmv", self._ctx, ctypes.c_void_p(x), ctypes.c_void_p(y))

    # -- linear buckling analysis (fcVM.py:1199-1212) ----------------------------------------------------------
    def set_coordinates(self, nocoord):
        """New nodal coordinates on the same topology (the imperfect geometry, fcVM.py:1240)."""
        xyz = _np(nocoord, np.float64)
        assert xyz.shape == self._nocoord.shape
        self._nocoord = xyz
        call("fcvm_set_coordinates", self._ctx, _ptr(xyz, f64p))

    def assemble_buckling(self, sigma):
        """K (not eliminated, prescribed diagonals x 100) and G = -nsm of the stress state in SIG_NEW; the
        engine's matrix becomes K - sigma G (``solve`` / ``spmv``), ``spmv_geometric`` multiplies with G."""
        call("fcvm_assemble_buckling", self._ctx, float(sigma))

    def spmv_geometric(self, x, y):
        call("fcvm_spmv_geometric", self._ctx, ctypes.c_void_p(x), ctypes.c_void_p(y))

    def buckling_modes(self, k=2, sigma=0.1, tol=1e-11, rtol=1e-12, max_sweeps=200, log=None):
        """The ``k`` eigenpairs of K x = lambda G x nearest ``sigma`` -- what the reference asks of ARPACK with
        ``eigsh(K, k=2, M=G, sigma=0.1, which='LM', mode='buckling')`` (fcVM.py:1212) -- by shift-invert subspace
        iteration on the device: Y = (K - sigma G)^-1 G X (one PCG solve per vector), Rayleigh-Ritz on span(Y)
        with the Gram matrices Y^T G Y and Y^T (K - sigma G) Y, until the Ritz values stand still.  Needs K - sigma G
        positive definite (sigma below the first buckling factor); otherwise the PCG reports the breakdown.
        Returns (eigenvalues ascending, eigenvectors as columns, unit 2-norm, largest component positive)."""
        import scipy.linalg as sla
        self.assemble_buckling(sigma)
        p = max(k + 2, 2 * k)
        rng = np.random.default_rng(12345)
        n = self.ndof
        X = [self.vec(host=rng.standard_normal(n)) for _ in range(p)]
        W = [self.vec() for _ in range(p)]
        Y = [self.vec() for _ in range(p)]
        GY = [self.vec() for _ in range(p)]
        theta_old = None
        lam = None
        for sweep in range(max_sweeps):
            for j in range(p):
                self.spmv_geometric(X[j], W[j])                    # W = G X = M Y
                self.solve(W[j], Y[j], rtol=rtol, max_iter=100000)
                self.spmv_geometric(Y[j], GY[j])
            a = np.empty((p, p))
            b = np.empty((p, p))
            for i in range(p):
                for j in range(i, p):
                    a[i, j] = a[j, i] = self.dot(Y[i], GY[j])
                    b[i, j] = b[j, i] = 0.5 * (self.dot(Y[i], W[j]) + self.dot(Y[j], W[i]))
            theta, cvec = sla.eigh(a, b)                           # a c = theta b c, b = Y^T M Y positive definite
            order = np.argsort(-np.abs(theta))                     # nearest sigma first
            theta, cvec = theta[order], cvec[:, order]
            for j in range(p):                                     # X = Y c
                self.axpby(cvec[0, j], Y[0], 0.0, X[j])
                for i in range(1, p):
                    self.axpby(cvec[i, j], Y[i], 1.0, X[j])
            lam = sigma + 1.0 / theta[:k]
            if log:
                log(f"buckling sweep {sweep}: load factors {np.sort(lam)}")
            if theta_old is not None and np.all(np.abs(theta[:k] - theta_old) <= tol * np.abs(theta[:k])):
                break
            theta_old = theta[:k].copy()
        vec = np.stack([self.get(X[j]) for j in range(k)], axis=1)
        for v in X + W + Y + GY:
            call("fcvm_vec_free", self._ctx, ctypes.c_void_p(v))
            self._vecs.remove(v)
        order = np.argsort(lam)
        lam, vec = lam[order], vec[:, order]
        for j in range(k):
            vec[:, j] /= np.linalg.norm(vec[:, j])
            if vec[np.argmax(np.abs(vec[:, j])), j] < 0:
                vec[:, j] = -vec[:, j]
        return lam, vec

    def matfree_apply(self, x, y):
        """y = K x with the elastic operator recomputed element by element (what the PCG uses for GNLN)."""
        call("fcvm_matfree_apply", self._ctx, ctypes.c_void_p(x), ctypes.c_void_p(y))

    def solve(self, b, x, rtol=1e-10, max_iter=20000, use_x0=False, raise_on_noconv=True, recycle=False):
        """x = K^-1 b by preconditioned CG.  Returns (iterations, relative residual).  ``recycle``: start from the
        projection onto the last two solutions of this matrix (kept in the library) and remember this one."""
        it = ctypes.c_int()
        rr = ctypes.c_double()
        rc = call("fcvm_pcg_solve", self._ctx, ctypes.c_void_p(b), ctypes.c_void_p(x), float(rtol), int(max_iter),
                  1 if use_x0 else (2 if recycle else 0), ctypes.byref(it), ctypes.byref(rr), allow=(_lib.E_NOCONV,))
        if rc == _lib.E_NOCONV and raise_on_noconv:
            raise FcvmError(rc, _lib.cdll().fcvm_last_error().decode())
        self.last_solve = (it.value, rr.value)
        self.pcg_iterations = getattr(self, "pcg_iterations", 0) + it.value
        self.pcg_solves = getattr(self, "pcg_solves", 0) + 1
        return it.value, rr.value

    def update_stress_load(self, disp_new, du, qin, Et_E, LD=False, yield_scale=1.0):
        call("fcvm_update_stress_load", self._ctx, ctypes.c_void_p(disp_new) if disp_new else None,
             ctypes.c_void_p(du), ctypes.c_void_p(qin), float(Et_E), 1 if LD else 0, float(yield_scale))

    def update_peeq_csr(self, ultimate_strain, Et_E):
        """Returns (argmax_gp, csr_max, pressure, sigmises, triax, ecr, peeq, peeq_max) -- fcVM.py:1543-1554."""
        arg = ctypes.c_int64()
        out = (ctypes.c_double * 7)()
        call("fcvm_update_peeq_csr", self._ctx, float(ultimate_strain), float(Et_E), ctypes.byref(arg), out)
        res = (arg.value,) + tuple(out)
        if self.comm is not None and self.comm.world > 1:
            # global Gauss-point number = local + 4 * first element of the rank; first maximum wins (np.argmax)
            mine = (res[0] + 4 * self.comm.elem_offset(),) + res[1:]
            every = self.comm.allgather(mine)
            best = max(every, key=lambda t: (t[1], -t[0]))
            res = best[:7] + (max(t[7] for t in every),)
        return res

    def scale_step_stress(self, fac):
        call("fcvm_scale_step_stress", self._ctx, float(fac))

    def map_stresses(self, averaged, sig_yield, noce=None):
        nn = self.nn
        t10s = np.empty((nn, 6))
        outs = [np.empty(nn) for _ in range(4)]
        nc = _np(noce, np.int16) if noce is not None else None
        call("fcvm_map_stresses", self._ctx, 1 if averaged else 0, float(sig_yield),
             _ptr(nc, i16p) if nc is not None else None, _ptr(t10s, f64p), *[_ptr(o, f64p) for o in outs])
        return (t10s, *outs)

    def interface_sum(self, v):
        call("fcvm_interface_sum", self._ctx, ctypes.c_void_p(v))

    def synchronize(self):
        call("fcvm_synchronize", self._ctx)

    # -- measurement ------------------------------------------------------------------------------
    def timer_start(self):
        call("fcvm_timer_start", self._ctx)

    def timer_stop_ms(self) -> float:
        ms = ctypes.c_float()
        call("fcvm_timer_stop_ms", self._ctx, ctypes.byref(ms))
        return ms.value

    def profile(self, on):
        """0/False off, 1/True every launch (synchronous), k >= 2 every k-th launch (asynchronous)."""
        call("fcvm_profile_enable", self._ctx, int(on))
        call("fcvm_profile_reset", self._ctx)

    def profile_get(self):
        """family -> (summed ms of timed launches, timed launches, all launches)."""
        names = ("spmv", "stress_update", "node_gather", "pcg_vector", "assembly", "elem_stiffness", "coo_reduce",
                 "peer_exchange", "coarse_rhs", "coarse_product", "coarse_expand", "pcg_step")
        out = {}
        for i, k in enumerate(names):
            ms = ctypes.c_double()
            n = ctypes.c_int64()
            call("fcvm_profile_get", self._ctx, i, ctypes.byref(ms), ctypes.byref(n))
            out[k] = (ms.value, n.value, int(call("fcvm_profile_seen", self._ctx, i)))
        return out

    def launch_count(self) -> int:
        return int(call("fcvm_launch_count", self._ctx))

    def copy_bytes(self):
        """(host->device, device->host) bytes this engine's context has moved so far, counted in the library."""
        a, b = ctypes.c_int64(), ctypes.c_int64()
        call("fcvm_copy_bytes", self._ctx, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def deflation_stats(self):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        call("fcvm_deflation_stats", self._ctx, ctypes.byref(a), ctypes.byref(b))
        return dict(boxes=a.value, entries=b.value)

    def matrix_stats(self):
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        call("fcvm_matrix_stats", self._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return dict(blocks_stored=a.value, blocks_real=b.value, bytes=c.value)

    # -- host-buffer drop-ins ------------------------------------------------------------------------
    def host_update_stress_load(self, sig_yield, disp_new, du, sig, sig_update, sig_test_global, qin, Et_E, LD, pgp):
        pg = np.zeros(4 * self.ne, dtype=np.uint8)
        dn = _np(disp_new, np.float64) if disp_new is not None else None
        call("fcvm_host_update_stress_load", self._ctx, _ptr(_np(sig_yield, np.float64), f64p),
             _ptr(dn, f64p) if dn is not None else None, _ptr(_np(du, np.float64), f64p),
             _ptr(_np(sig, np.float64), f64p), _ptr(sig_update, f64p), _ptr(sig_test_global, f64p), _ptr(qin, f64p),
             float(Et_E), 1 if LD else 0, _ptr(pg, u8p))
        pgp[:] = pg.astype(bool)

    def host_solve(self, b, rtol=1e-10, max_iter=20000, out=None, raise_on_noconv=True, recycle=False):
        x = np.empty(self.ndof) if out is None else out
        it = ctypes.c_int()
        rr = ctypes.c_double()
        rc = call("fcvm_host_solve", self._ctx, _ptr(_np(b, np.float64), f64p), _ptr(x, f64p), float(rtol),
                  int(max_iter), 1 if recycle else 0, ctypes.byref(it), ctypes.byref(rr), allow=(_lib.E_NOCONV,))
        if rc == _lib.E_NOCONV and raise_on_noconv:
            raise FcvmError(rc, _lib.cdll().fcvm_last_error().decode())
        self.last_solve = (it.value, rr.value)
        self.pcg_iterations = getattr(self, "pcg_iterations", 0) + it.value
        self.pcg_solves = getattr(self, "pcg_solves", 0) + 1
        return x


def deflation_boxes(lo, hi, elem_extent, target_unknowns=3072, grid=None):
    """Box grid of the deflation level: about ``target_unknowns / 6`` near-cubic boxes over the bounding
    box ``lo..hi``, every box at least two elements wide (``elem_extent`` = widest element per direction:
    the nodes a node couples to then lie in at most 2 x 2 x 2 boxes) and at most 16384 coarse unknowns.
    Returns (boxes per direction, box size)."""
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    size = np.maximum(hi - lo, 1e-300)
    if grid is None:
        m = max(1, int(target_unknowns) // 6)
        h0 = (float(np.prod(size)) / m) ** (1.0 / 3.0)
        grid = np.maximum(1, np.round(size / h0).astype(int))
    grid = np.asarray(grid, dtype=int).copy()
    widest = np.floor(size / np.maximum(2.0001 * np.asarray(elem_extent, dtype=np.float64), 1e-300)).astype(int)
    grid = np.maximum(1, np.minimum(grid, widest))
    while 6 * int(np.prod(grid)) > 16384:
        grid[int(np.argmax(grid))] -= 1
    return grid, size / grid


# ----------------------------------------------------------------------------------------------
# Reference-signature functions (numpy in, numpy out).  An Engine is cached per mesh so that
# repeated calls -- the reference calls update_stress_load once per Newton iteration -- reuse
# the resident connectivity and sparsity pattern.
# ----------------------------------------------------------------------------------------------
_engine_cache = {}          # insertion-ordered: the oldest engine is evicted first


def _mesh_key(el, xyz, mat, device):
    """Identity of a mesh for the engine cache.  The reference hands the same arrays to update_stress_load on every
    Newton iteration, so the key must be cheap: shapes, material, and a strided sample of connectivity and
    coordinates (a few thousand values whatever the mesh size) instead of a hash over ~100 MB."""
    se = el.reshape(-1)[::max(1, el.size // 4096)]
    sx = xyz.reshape(-1)[::max(1, xyz.size // 4096)]
    return (el.shape, xyz.shape, hash(se.tobytes()), hash(sx.tobytes()), float(el.sum()), float(xyz.sum()),
            float(mat[0][0]), float(mat[0][1]), float(mat[0][2]), device)


def _engine_for(elNodes, nocoord, materialbyElement, fix=None, device=0) -> Engine:
    el = np.asarray(elNodes)
    xyz = np.asarray(nocoord)
    mat = np.asarray(materialbyElement, dtype=np.float64)
    key = _mesh_key(el, xyz, mat, device)
    eng = _engine_cache.get(key)
    if eng is None:
        if len(_engine_cache) >= 4:
            oldest = next(iter(_engine_cache))
            _engine_cache.pop(oldest).close()
        eng = Engine(el, xyz, mat, device=device)
        _engine_cache[key] = eng
    if fix is not None:
        eng.set_constraints(dict(fix))
    return eng


def clear_engine_cache():
    for e in _engine_cache.values():
        e.close()
    _engine_cache.clear()


def calcGSM(elNodes, nocoord, materialbyElement, fix, grav_x, grav_y, grav_z, loadfaces, pressure, loadvertices,
            vertexloads, loadedges, edgeloads, loadfaces_uni, faceloads, device=0):
    """Drop-in for fcVM.py:620-816.  Returns the reference's tuple; the stiffness comes back as
    the COO triplets of the assembled lower triangle (already summed -- scipy's csc_matrix of it
    equals the reference's ``gsm`` entry for entry)."""
    eng = _engine_for(elNodes, nocoord, materialbyElement, fix, device)
    glv_h = surface_load_vector(nocoord, loadfaces, pressure, loadvertices, vertexloads, loadedges, edgeloads,
                                loadfaces_uni, faceloads)
    glv = eng.vec(host=glv_h)
    eng.assemble(glv, (grav_x, grav_y, grav_z))
    indptr, indices, data = eng.export_csc_lower()
    col = np.repeat(np.arange(eng.ndof, dtype=np.int64), np.diff(indptr))
    glv_h = eng.get(glv)
    modf = eng.get(eng.buf(MODF))
    call("fcvm_vec_free", eng._ctx, ctypes.c_void_p(glv))
    eng._vecs.remove(glv)
    ls = glv_h.reshape(-1, 3).sum(axis=0)
    x = gauss_point_coordinates(elNodes, nocoord)
    V = float("nan")          # "Element volume - not used" (fcVM.py:760)
    return data, indices, col, glv_h, modf, V, ls[0], ls[1], ls[2], eng.ne, eng.nn, x


def gauss_point_coordinates(elNodes, nocoord, gps=None):
    """``x`` of calcGSM (fcVM.py:761): coordinates of the Gauss points, (4*ne, 3) or selected rows."""
    a, b = 0.138196601125011, 0.585410196624968
    pts = np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]])
    xi, et, ze = pts[:, 0], pts[:, 1], pts[:, 2]
    c = 1.0 - xi - et - ze
    shp = np.stack([(2 * c - 1) * c, xi * (2 * xi - 1), et * (2 * et - 1), ze * (2 * ze - 1), 4 * xi * c, 4 * xi * et,
                    4 * et * c, 4 * ze * c, 4 * xi * ze, 4 * et * ze], axis=1)           # (4, 10)
    el = np.asarray(elNodes)
    xyz = np.asarray(nocoord, dtype=np.float64)
    if gps is None:
        return np.einsum("gk,ekc->egc", shp, xyz[el - 1]).reshape(-1, 3)
    gps = np.asarray(gps, dtype=np.int64)
    return np.einsum("nk,nkc->nc", shp[gps % 4], xyz[el[gps // 4] - 1])


def update_stress_load(gp10, elNodes, nocoord, materialbyElement, sig_yield, disp_new, du, sig, sig_update,
                       sig_test_global, qin, Et_E, LD, pgp, device=0):
    """Drop-in for fcVM.py:2196-2464 (host arrays in and out, arithmetic on the GPU)."""
    eng = _engine_for(elNodes, nocoord, materialbyElement, None, device)
    eng.host_update_stress_load(sig_yield, disp_new, du, sig, sig_update, sig_test_global, qin, Et_E, LD, pgp)


def update_PEEQ_CSR(nelem, materialbyElement, sig_test, sig_new, sig_yield, ultimate_strain, peeq, csr, triax,
                    pressure, sigmises, ecr, Et_E, engine: Optional[Engine] = None):
    """Drop-in for fcVM.py:2084-2137; needs the mesh's Engine (the arrays alone carry no mesh)."""
    if engine is None:
        raise FcvmError(-1, "update_PEEQ_CSR: pass engine=Engine(...) for the mesh these arrays belong to")
    e = engine
    e.gp_put(SIG_TEST, sig_test)
    e.gp_put(SIG_NEW, sig_new)
    for which, arr in ((SIG_YIELD, sig_yield), (PEEQ, peeq), (CSR, csr)):
        e.gp_put(which, arr)
    res = e.update_peeq_csr(ultimate_strain, Et_E)
    for which, arr in ((SIG_YIELD, sig_yield), (PEEQ, peeq), (CSR, csr), (TRIAX, triax), (PRESSURE, pressure),
                       (SIGMISES, sigmises), (ECR, ecr)):
        arr[:] = e.gp_get(which)
    return res


def mapStresses(averaged, elNodes, nocoord, sig, peeq, sigvm, csr, noce, sig_yield, device=0):
    """Drop-in for fcVM.py:2496-2554."""
    eng = _engine_for(elNodes, nocoord, np.array([[1.0, 0.3, 0.0]]), None, device)
    eng.gp_put(SIG_NEW, sig)
    eng.gp_put(PEEQ, peeq)
    eng.gp_put(SIGMISES, sigvm)
    eng.gp_put(CSR, csr)
    return eng.map_stresses(averaged, sig_yield, noce)


# ----------------------------------------------------------------------------------------------
# calcDisp: the load-stepping driver (fcVM.py:1083-1635), vectors resident on the device.
# ----------------------------------------------------------------------------------------------
def calcDisp(model, ctl, clicks=(), device=0, rtol=1e-10, max_iter=50000, log=None, engine: Optional[Engine] = None,
             comm=None, on_iteration=None, deflation="auto"):
    """Run the whole load-displacement analysis on the GPU.

    Same control flow as the reference (arc-length corrected modified Newton with restarts,
    fcVM.py:1304-1559); ``clicks`` scripts the interactive window ("add", "rev", "stop",
    ("add", target)).  The direct CHOLMOD solves are replaced by PCG to ``rtol``.  Returns a
    dictionary with the reference's return values plus iteration statistics.
    """
    m = model
    say = log or (lambda *a: None)
    own = engine is None
    eng = engine or Engine(m.elNodes, m.nocoord, m.materialbyElement, m.fix, device=device, comm=comm)
    if not own:
        eng.set_constraints(m.fix)
    if deflation == "auto":
        # second preconditioner level (rigid-body-mode deflation) once the mesh is large enough
        # for boxes two elements wide to be worth it; an engine handed in keeps whatever its owner chose
        deflation = (AUTO_DEFLATION if eng.nn * (comm.world if comm is not None else 1) >= 20000 else 0) if own else None
    if deflation is not None and hasattr(eng, "set_deflation"):
        eng.set_deflation(deflation)
    ndof, nelem = eng.ndof, eng.ne
    nstep, iterat_max, error_max = ctl.nstep, ctl.iterat_max, ctl.error_max
    relax, scale_re, scale_up, scale_dn = ctl.relax, ctl.scale_re, ctl.scale_up, ctl.scale_dn
    disp_output, ultimate_strain, Et_E, target_LF = ctl.disp_output, ctl.ultimate_strain, ctl.Et_E, ctl.target_LF
    gnl, maxImp = ctl.gnl, float(ctl.maxImp)
    grav = (ctl.grav_x, ctl.grav_y, ctl.grav_z)
    if gnl == "GNLY":                                              # fcVM.py:1087-1097
        LD, relax, disp_output, scale_up = True, 1.0, "total", 1.1
    else:
        LD = False
    buckling = LD and not (float(nstep) > 1.0 and maxImp == 0.0)  # fcVM.py:1200
    if buckling and comm is not None and comm.world > 1:
        raise NotImplementedError("the linear buckling pre-analysis runs on one GPU")

    coords = [m.nocoord]            # replaced by the imperfect geometry after a buckling pre-analysis

    def load_vector(disp_host=None):
        return surface_load_vector(coords[0], m.loadfaces, m.pressure, m.loadvertices, m.vertexloads, m.loadedges,
                                   m.edgeloads, m.loadfaces_uni, m.faceloads, disp=disp_host)

    # follower loads: only surface / edge / vertex loads depend on the displaced geometry (calcTSM, fcVM.py:856-938);
    # without any the large-displacement branch keeps its load vector on the device
    has_loads = any(len(t) > 1 for t in (m.loadfaces, m.loadfaces_uni, m.loadedges, m.loadvertices))
    glv = eng.vec(host=load_vector())
    eng.assemble(glv, grav)                                        # calcGSM
    modf, fixdof = eng.buf(MODF), eng.buf(FIXDOF)
    movdof_any = bool(np.max(m.movdof) == 1)
    cm = getattr(eng, "comm", None)
    if cm is not None and cm.world > 1:
        # every rank must take the same branches (they contain collectives): the flag is global
        movdof_any = any(cm.allgather(movdof_any))
    loadsum = eng.get(glv).reshape(-1, 3).sum(axis=0)

    qnorm = eng.norm(glv)                                          # fcVM.py:1115-1116
    if qnorm < 1.0:
        qnorm = 1.0
    V = eng.vec
    f, ue, du, a, due, r, qin = V(), V(), V(), V(), V(), V(), V()
    disp_new, disp_old, zero = V(), V(), V()
    # f = fixdof*glv + modf ; ue = K^-1 f                          fcVM.py:1128-1135
    eng.residual(1.0, glv, zero, f)                                # f = fixdof*(1*glv - 0)
    eng.axpby(1.0, modf, 1.0, f)
    eng.solve(f, ue, rtol, max_iter)
    disp_el = eng.get(ue)

    dl0 = 1.0 / nstep
    dl = dl0
    eng.axpby(dl, ue, 0.0, du)                                     # du = dl*ue
    eng.gp_fill(SIG_YIELD, ctl.sig_yield)

    if movdof_any:                                                 # fcVM.py:1169-1177
        eng.update_stress_load(disp_new, ue, qin, Et_E, LD)
        qnorm = eng.masked_norm(qin, m.movdof)                     # ||movdof * qelastic||

    eigenval, eigenvec = None, None
    nocoord_old = np.array(m.nocoord)
    if buckling:
        # linear buckling analysis (fcVM.py:1195-1214): the elastic stress state of the full load, K and G of
        # calcTSM's nstep == 1 branch, the two load factors nearest 0.1
        eng.gp_fill(SIG_OLD, 0.0)
        eng.update_stress_load(zero, ue, qin, Et_E, False, yield_scale=1.0e6)
        eigenval, eigenvec = eng.buckling_modes(k=2, sigma=0.1, log=say)
        say(f"buckling load factors: {eigenval}")
        if float(nstep) != 1.0 and maxImp != 0.0:                  # imperfection and restart, fcVM.py:1224-1294
            ev1, ev2 = float(ctl.ev1), float(ctl.ev2)
            ua = ev1 / (ev1 + ev2) * eigenvec[:, 0] + ev2 / (ev1 + ev2) * eigenvec[:, 1]
            ub = ev1 / (ev1 + ev2) * eigenvec[:, 0] - ev2 / (ev1 + ev2) * eigenvec[:, 1]
            ma, mb = np.max(np.abs(ua)), np.max(np.abs(ub))
            if ma > mb:
                imper = maxImp / ma * np.sign(ua[np.argmax(np.abs(ua))]) * ua
            else:
                imper = maxImp / mb * np.sign(ub[np.argmax(np.abs(ub))]) * ub
            coords[0] = m.nocoord + imper.reshape(-1, 3)           # the imperfect geometry from here on
            eng.set_coordinates(coords[0])
            eng.put(glv, load_vector())
            # back to the eliminated elastic operator, now of the imperfect geometry: calcGSM, factorisation and
            # elastic solution again, state reset (fcVM.py:1242-1294)
            if deflation is not None and hasattr(eng, "set_deflation"):
                eng.set_deflation(deflation)
            eng.assemble(glv, grav)
            qnorm = eng.norm(glv)
            if qnorm < 1.0:
                qnorm = 1.0
            eng.residual(1.0, glv, zero, f)
            eng.axpby(1.0, modf, 1.0, f)
            eng.solve(f, ue, rtol, max_iter)
            disp_el = eng.get(ue)
            dl = dl0
            eng.axpby(dl, ue, 0.0, du)
    step = -1
    cnt = True
    fail = False
    un, csrplot, crip, pplot, svmplot, triaxplot, peeqplot, peeqmax, ecrplot = (
        [0.], [0.], [0], [0.], [0.], [0.], [0.], [0.], [0.])
    lbd = [0.0]
    rfl = [0.0]
    iters, nplastic, pcg_its = [], [], []
    for b in (SIG_NEW, SIG_OLD, SIG_TEST):
        eng.gp_fill(b, 0.0)
    if float(nstep) == 1.0:
        # elastic analysis (fcVM.py:1216-1223): displacements of the elastic solve, no load stepping.  The
        # reference wipes sig_new after its elastic stress call (fcVM.py:1195-1197, then 1300), so the
        # returned stresses are zero there and here.
        eng.copy(ue, disp_new)
        lbd.append(1.0)
        rfl.append(1.0)
        un.append(float(np.max(np.abs(disp_el))))
        cnt = False
    iterat_tot = 0
    mrr = False
    queue = list(clicks)
    aa = 0.0
    lout = [0.0] if float(nstep) == 1.0 else lbd              # fcVM.py:1193: never reassigned without load steps

    def record():
        res = eng.update_peeq_csr(ultimate_strain, Et_E)
        crip.append(int(res[0]))
        csrplot.append(res[1])
        pplot.append(res[2])
        svmplot.append(res[3])
        triaxplot.append(res[4])
        ecrplot.append(res[5])
        peeqplot.append(res[6])
        peeqmax.append(res[7])

    def stress_update():
        eng.update_stress_load(disp_new, du, qin, Et_E, LD)

    stopped = False
    try:
        while cnt:
            cnt = False
            pstep = 0
            while pstep < nstep and not mrr:
                step += 1
                pstep += 1
                restart = 0
                say(f"Step: {step}")
                eng.copy(du, a)                                        # a: Riks control vector
                eng.gp_copy(SIG_NEW, SIG_OLD)
                lbd.append(lbd[step] + dl)
                stress_update()
                rnorm = eng.residual(lbd[step + 1], glv, qin, r)
                error = rnorm / qnorm
                iterat = 0
                say(f"Iteration: {iterat}, Error: {error:.2e}")
                while error > error_max and not mrr:
                    iterat += 1
                    iterat_tot += 1
                    if LD and (iterat == 1 or eng.plastic_count() > 0):       # fcVM.py:1351-1396
                        if has_loads:
                            eng.put(glv, load_vector(eng.get(disp_new)))
                        else:
                            eng.zero(glv)                                         # gravity is added by the assembly
                        eng.assemble(glv, grav, tangent=True, disp=disp_new, Et_E=Et_E)
                        eng.residual(1.0, glv, zero, f)
                        eng.axpby(1.0, modf, 1.0, f)
                        eng.solve(f, ue, rtol, max_iter)
                        eng.copy(ue, a)
                        eng.axpby(0.0, a, eng.norm(du) / eng.norm(a), a)      # a *= |du|/|a|
                    eng.axpby(relax, r, 0.0, f)                               # f = relax*r
                    its, _ = eng.solve(f, due, rtol, max_iter, recycle=not LD)
                    pcg_its.append(its)
                    dl = -eng.dot(a, due) / eng.dot(a, ue)                    # Riks correction, fcVM.py:1414-1417
                    lbd[step + 1] += dl
                    aa = eng.norm(a)
                    eng.axpbypcz(1.0, due, dl, ue, 1.0, du)                   # du += due + dl*ue
                    uu = eng.norm(du)
                    sf = min(aa / uu, 1.0)
                    lbd[step + 1] = lbd[step] + sf * (lbd[step + 1] - lbd[step])
                    eng.axpby(0.0, du, sf, du)                                # du *= sf
                    stress_update()
                    rnorm = eng.residual(lbd[step + 1], glv, qin, r)
                    error = rnorm / qnorm
                    say(f"Iteration: {iterat}, Error: {error:.2e}")
                    if on_iteration is not None:
                        on_iteration(dict(eng=eng, step=step, iterat=iterat, iterat_tot=iterat_tot, error=error,
                                          pcg_iterations=its, lbd=lbd))
                    if iterat > iterat_max:                                   # fcVM.py:1457-1484
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
                            eng.axpbypcz(1.0 / scale_re / restart, disp_new, -1.0 / scale_re / restart, disp_old, 0.0, du)
                        elif not mrr:
                            dl = dl0 / scale_re / restart
                            eng.axpby(dl / scale_re / restart, ue, 0.0, du)
                        # unconditional in the reference (fcVM.py:1474): after MAXIMUM RESTARTS the last kept load
                        # level is overwritten with lbd[step] + the last Riks dl
                        lbd[step + 1] = lbd[step] + dl
                        if not mrr:
                            stress_update()
                            # r = fixdof*(lbd*(glv+modf) - qin): fixdof*modf is modf on free dofs
                            eng.axpbypcz(1.0, glv, 1.0, modf, 0.0, f)
                            rnorm = eng.residual(lbd[step + 1], f, qin, r)
                            error = rnorm / qnorm
                            iterat = 0
                if abs(target_LF - lbd[step]) < abs(lbd[step + 1] - lbd[step]):           # fcVM.py:1486-1510
                    say("REACHED TARGET LOAD")
                    fac = (target_LF - lbd[step]) / (lbd[step + 1] - lbd[step])
                    eng.axpby(0.0, du, fac, du)
                    eng.scale_step_stress(fac)
                    lbd[step + 1] = target_LF
                    eng.axpby(1.0, du, 1.0, disp_new)
                    un.append(eng.max_node_disp(disp_new))
                    record()
                    iters.append(iterat)
                    nplastic.append(eng.plastic_count())
                    break
                elif not mrr:                                                                 # fcVM.py:1515-1559
                    eng.copy(disp_new, disp_old)
                    eng.axpby(1.0, du, 1.0, disp_new)
                    dl = lbd[step + 1] - lbd[step]
                    if movdof_any:
                        rfl.append(eng.reaction(qin))
                    if iterat > 10:
                        dl /= scale_dn
                        eng.axpby(0.0, du, 1.0 / scale_dn, du)
                    if iterat < 5:
                        dl *= scale_up
                        eng.axpby(0.0, du, scale_up, du)
                    un.append(eng.max_node_disp(disp_new))
                    record()
                    iters.append(iterat)
                    nplastic.append(eng.plastic_count())
            lout = rfl if movdof_any else lbd
            if queue and not mrr:                       # scripted stand-in for the plot window (fcVM.py:1639-2080)
                ev = queue.pop(0)
                tgt = target_LF
                if isinstance(ev, tuple):
                    ev, tgt = ev
                if ev == "add":
                    LF = lout[-1]
                    if (target_LF - LF) * (tgt - LF) <= 0.0:
                        dl = float(np.sign(tgt - LF)) * 1.0 / nstep
                        eng.axpby(dl, ue, 0.0, du)
                    cnt = True
                elif ev == "rev":
                    dl = -dl
                    eng.axpby(0.0, du, -1.0, du)
                    cnt = True
                target_LF = tgt

    except StopAnalysis:
        stopped = True
        lout = rfl if movdof_any else lbd

    dn = eng.get(disp_new)
    dis = dn if disp_output == "total" else dn - eng.get(disp_old)
    crip_a = np.asarray(crip)
    out = dict(displacements=dis, disp_el=disp_el, stresses=eng.gp_get(SIG_NEW), peeq=eng.gp_get(PEEQ),
               sigmises=eng.gp_get(SIGMISES), csr=eng.gp_get(CSR), lout=np.asarray(lout), un=np.asarray(un),
               crip=crip_a, peeqplot=np.asarray(peeqplot), pplot=np.asarray(pplot), svmplot=np.asarray(svmplot),
               triaxplot=np.asarray(triaxplot), ecrplot=np.asarray(ecrplot), csrplot=np.asarray(csrplot), fail=fail,
               nocoord_old=nocoord_old, eigenval=eigenval, eigenvec=eigenvec, lbd=np.asarray(lbd), iters=np.asarray(iters),
               nplastic=np.asarray(nplastic), iterat_tot=iterat_tot, pcg_iterations=np.asarray(pcg_its),
               glv=eng.get(glv), modf=eng.get(modf), loadsum=tuple(loadsum), sig_yield=eng.gp_get(SIG_YIELD),
               pgp=eng.gp_get(PGP), sig_test=eng.gp_get(SIG_TEST),
               x_crip=(gauss_point_coordinates(m.elNodes, m.nocoord, crip_a)
                       if getattr(eng, "comm", None) is None or eng.comm.world == 1 else None), ne=eng.ne, nn=eng.nn,
               launches=eng.launch_count(), stopped=stopped)
    if own:
        eng.close()
    return out
