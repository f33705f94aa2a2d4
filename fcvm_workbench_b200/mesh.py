"""Synthetic structured C3D10 meshes (the benchmark workloads of BASELINE.json).

The reference takes its second-order tetrahedra from Gmsh/Netgen through
FreeCAD (source code/fcVM.py:136-164).  For the throughput runs we need meshes
of a chosen size without a mesher: a box of ``nx*ny*nz`` cells, each cut into
six tetrahedra around the cell diagonal (Kuhn / Freudenthal subdivision, which
is conforming across cells), with every lattice point of the doubled grid a
node.  Node numbers are 1-based and the local node order is the CalculiX one
the element routines expect after ``setUpInput``'s swap (fcVM.py:338-341):
corners 0-3, then mid-side nodes (0,1) (1,2) (0,2) (0,3) (1,3) (2,3).
"""
from __future__ import annotations

import itertools

import numpy as np

from .model import Model, count_noce, empty_loads, finish_bcs

_MID_PAIRS = ((0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3))


def _kuhn_corners():
    """Corner offsets (in cell units) of the six tetrahedra, positively oriented."""
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [np.zeros(3, dtype=np.int64)]
        for ax in perm:
            nxt = v[-1].copy()
            nxt[ax] += 1
            v.append(nxt)
        v = np.array(v)
        vol = np.linalg.det((v[1:] - v[0]).astype(float))
        if vol < 0:
            v[[1, 2]] = v[[2, 1]]
        tets.append(v)
    return np.array(tets)            # (6, 4, 3)


def box_mesh(nx: int, ny: int, nz: int, lx: float = 1.0, ly: float = 1.0, lz: float = 1.0):
    """``(elNodes, nocoord)`` of the structured box, both in the reference's conventions."""
    mx, my, mz = 2 * nx + 1, 2 * ny + 1, 2 * nz + 1
    gi, gj, gk = np.meshgrid(np.arange(mx), np.arange(my), np.arange(mz), indexing="ij")
    # x fastest: id = 1 + i + mx*(j + my*k)
    nocoord = np.empty((mx * my * mz, 3))
    order = (gi + mx * (gj + my * gk)).ravel()
    nocoord[order, 0] = gi.ravel() * (lx / (2 * nx))
    nocoord[order, 1] = gj.ravel() * (ly / (2 * ny))
    nocoord[order, 2] = gk.ravel() * (lz / (2 * nz))

    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    # cells ordered x fastest as well, so that element and node numbers stay local together
    cell = np.stack([ci.transpose(2, 1, 0).ravel(), cj.transpose(2, 1, 0).ravel(), ck.transpose(2, 1, 0).ravel()],
                    axis=1)                                     # (nc, 3)
    tets = _kuhn_corners()                                      # (6, 4, 3)
    corner = 2 * (cell[:, None, None, :] + tets[None, :, :, :])  # (nc, 6, 4, 3) lattice coordinates
    corner = corner.reshape(-1, 4, 3)
    lat = np.empty((corner.shape[0], 10, 3), dtype=np.int64)
    lat[:, :4] = corner
    for m, (a, b) in enumerate(_MID_PAIRS):
        lat[:, 4 + m] = (corner[:, a] + corner[:, b]) // 2
    elNodes = 1 + lat[..., 0] + mx * (lat[..., 1] + my * lat[..., 2])
    return elNodes.astype(np.int64), nocoord


def cube_model(n: int, size: float = 10.0, mode: str = "platen", top_disp: float = 1.0, E: float = 210000.0,
               nu: float = 0.3, density: float = 7.85e-6, name: str | None = None, nxyz=None) -> Model:
    """Displacement-controlled block of ``6 n^3`` C3D10 elements.

    ``mode``
      * ``"tension"``  - symmetry planes x=0, y=0, z=0 and a prescribed ``uz`` on the top
        face (homogeneous uniaxial stress; the closed-form check).
      * ``"platen"``   - bottom face clamped, top face moved by ``uz`` with ``ux=uy=0``
        (rough rigid platen: stress concentrations, mixed elastic/plastic Gauss points).
      * ``"punch"``    - bottom clamped, the strip ``x <= size/4`` of the top face pushed
        down (footing-type collapse mechanism).
      * ``"force"``    - like ``"tension"`` but load-controlled by a uniform top traction of
        ``top_disp`` N/mm^2 (exercises glv / loadfaces_uni).
    """
    nx, ny, nz = nxyz if nxyz is not None else (n, n, n)
    elNodes, nocoord = box_mesh(nx, ny, nz, size, size * ny / nx, size * nz / nx)
    nn = len(nocoord)
    tol = 1e-9 * size
    top_z = nocoord[:, 2].max()
    ids = np.arange(1, nn + 1)
    bottom = ids[np.abs(nocoord[:, 2]) < tol]
    top = ids[np.abs(nocoord[:, 2] - top_z) < tol]
    x0 = ids[np.abs(nocoord[:, 0]) < tol]
    y0 = ids[np.abs(nocoord[:, 1]) < tol]
    loads = empty_loads()
    T, F = True, False
    if mode == "tension":
        disp = [(bottom, [T, T, F], [0, 0, 0]), (x0, [F, T, T], [0, 0, 0]), (y0, [T, F, T], [0, 0, 0]),
                (top, [T, T, F], [0, 0, top_disp])]
    elif mode == "platen":
        disp = [(bottom, [F, F, F], [0, 0, 0]), (top, [F, F, F], [0, 0, top_disp])]
    elif mode == "punch":
        strip = ids[(np.abs(nocoord[:, 2] - top_z) < tol) & (nocoord[:, 0] <= 0.25 * size + tol)]
        disp = [(bottom, [F, F, F], [0, 0, 0]), (strip, [T, T, F], [0, 0, -abs(top_disp)])]
    elif mode == "force":
        disp = [(bottom, [T, T, F], [0, 0, 0]), (x0, [F, T, T], [0, 0, 0]), (y0, [T, F, T], [0, 0, 0])]
        faces = top_faces(elNodes, nocoord, top_z, tol)
        loads["loadfaces_uni"] = np.vstack([loads["loadfaces_uni"], faces])
        loads["faceloads"] = np.vstack([loads["faceloads"], np.tile([0.0, 0.0, top_disp], (len(faces), 1))])
    else:
        raise ValueError(mode)
    fix, fixdof, movdof = finish_bcs(nn, disp)
    mat = np.tile(np.array([E, nu, density]), (len(elNodes), 1))
    return Model(name=name or f"cube{n}_{mode}", elNodes=elNodes, nocoord=nocoord, fix=fix, fixdof=fixdof,
                 movdof=movdof, materialbyElement=mat, noce=count_noce(elNodes, nn), **loads)


_TET_FACES = ((0, 1, 2, 4, 5, 6), (0, 1, 3, 4, 8, 7), (1, 2, 3, 5, 9, 8), (0, 2, 3, 6, 9, 7))


def top_faces(elNodes, nocoord, z, tol):
    """Six-node triangles (corners, then mid-sides (0,1) (1,2) (2,0)) lying in the plane ``z``."""
    out = []
    for f in _TET_FACES:
        nodes = elNodes[:, list(f)]
        on = np.all(np.abs(nocoord[nodes - 1, 2] - z) < tol, axis=1)
        out.append(nodes[on])
    return np.vstack(out)


def plate_with_hole_model(nr: int = 4, nt: int = 8, nz: int = 1, L: float = 50.0, R: float = 10.0, T: float = 5.0,
                          pull: float = 0.2, E: float = 210000.0, nu: float = 0.3, density: float = 7.85e-6,
                          grading: float = 1.6) -> Model:
    """Quarter of a square plate with a central circular hole, pulled in x (the analogue of the
    reference's ``Plate_with_hole_Example``: stress concentration at the hole, mixed elastic and
    plastic Gauss points).  The structured box mesh is mapped onto the region between the hole and
    the outer edges; mid-side nodes follow the map, so the elements next to the hole have curved
    edges (non-constant Jacobians).  Symmetry on x=0, y=0, z=0; ``ux = pull`` prescribed on x=L.
    """
    elNodes, lat = box_mesh(nr, nt, nz, 1.0, 1.0, 1.0)          # logical (s, t, zeta) in [0,1]^3
    s = lat[:, 0] ** grading                                     # finer towards the hole
    th = lat[:, 1] * (np.pi / 2)
    cx, cy = np.cos(th), np.sin(th)
    ox = np.where(th <= np.pi / 4, L, L * cx / np.maximum(cy, 1e-300))
    oy = np.where(th <= np.pi / 4, L * cy / np.maximum(cx, 1e-300), L)
    nocoord = np.stack([(1 - s) * R * cx + s * ox, (1 - s) * R * cy + s * oy, lat[:, 2] * T], axis=1)
    nn = len(nocoord)
    ids = np.arange(1, nn + 1)
    tol = 1e-9 * L
    x0 = ids[np.abs(nocoord[:, 0]) < tol]
    y0 = ids[np.abs(nocoord[:, 1]) < tol]
    z0 = ids[np.abs(nocoord[:, 2]) < tol]
    xL = ids[np.abs(nocoord[:, 0] - L) < tol]
    T_, F_ = True, False
    disp = [(z0, [T_, T_, F_], [0, 0, 0]), (x0, [F_, T_, T_], [0, 0, 0]), (y0, [T_, F_, T_], [0, 0, 0]),
            (xL, [F_, T_, T_], [pull, 0, 0])]
    fix, fixdof, movdof = finish_bcs(nn, disp)
    mat = np.tile(np.array([E, nu, density]), (len(elNodes), 1))
    return Model(name=f"plate_with_hole_{nr}x{nt}x{nz}", elNodes=elNodes, nocoord=nocoord, fix=fix, fixdof=fixdof,
                 movdof=movdof, materialbyElement=mat, noce=count_noce(elNodes, nn), **empty_loads())
