"""External load vector of the surface, edge and vertex loads (host side, O(surface)).

Reference: the first part of ``calcGSM`` (source code/fcVM.py:647-727) and of
``calcTSM`` (fcVM.py:856-938).  The tables carry the reference's dummy first row.
The volume part (gravity) is integrated on the device with the element
matrices (``fcvm_assemble``).
"""
from __future__ import annotations

import numpy as np

# 6-node triangle and 3-node line Gauss points (fcVM.py:598-611)
_GP6 = np.array([[0.445948490915965, 0.445948490915965, 0.111690794839005],
                 [0.10810301816807, 0.445948490915965, 0.111690794839005],
                 [0.445948490915965, 0.10810301816807, 0.111690794839005],
                 [0.091576213509771, 0.091576213509771, 0.054975871827661],
                 [0.816847572980458, 0.091576213509771, 0.054975871827661],
                 [0.091576213509771, 0.816847572980458, 0.054975871827661]])
_GP2 = np.array([[-0.5773502691896257, 1.0], [0.5773502691896257, 1.0]])


def _tri6(xi, et):
    """Shape functions and local derivatives of the 6-node triangle (fcVM.py:491-512)."""
    shp = np.array([(1.0 - xi - et) * (1.0 - 2.0 * xi - 2.0 * et), xi * (2.0 * xi - 1.0), et * (2.0 * et - 1.0),
                    4.0 * xi * (1.0 - xi - et), 4.0 * xi * et, 4.0 * et * (1 - xi - et)])
    d = np.array([[-3.0 + 4.0 * et + 4.0 * xi, -1.0 + 4.0 * xi, 0.0, -4.0 * (-1.0 + et + 2.0 * xi), 4.0 * et,
                   -4.0 * et],
                  [-3.0 + 4.0 * et + 4.0 * xi, 0.0, -1.0 + 4.0 * et, -4.0 * xi, 4.0 * xi,
                   -4.0 * (-1.0 + 2.0 * et + xi)]])
    return shp, d


def _face_integrals(xyz):
    """For faces given by node coordinates (nf, 6, 3): per Gauss point the weights
    ``N_i * |J| * w`` (nf, 6gp, 6) and the unit normals (nf, 6gp, 3)."""
    nf = xyz.shape[0]
    wts = np.empty((nf, 6, 6))
    nrm = np.empty((nf, 6, 3))
    for g, (xi, et, w) in enumerate(_GP6):
        shp, d = _tri6(xi, et)
        xs0 = np.einsum("k,fkc->fc", d[0], xyz)
        xs1 = np.einsum("k,fkc->fc", d[1], xyz)
        xp = np.cross(xs0, xs1)
        xsj = np.linalg.norm(xp, axis=1)
        nrm[:, g, :] = xp / xsj[:, None]
        wts[:, g, :] = shp[None, :] * (np.abs(xsj) * w)[:, None]
    return wts, nrm


def surface_load_vector(nocoord, loadfaces, pressure, loadvertices, vertexloads, loadedges, edgeloads,
                        loadfaces_uni, faceloads, disp=None):
    nocoord = np.asarray(nocoord, dtype=np.float64)
    nn = len(nocoord)
    glv = np.zeros(3 * nn)
    g3 = glv.reshape(nn, 3)
    pressure = np.asarray(pressure, dtype=np.float64)
    if len(pressure) > 1:
        nodes = np.asarray(loadfaces)[1:] - 1
        xyz = nocoord[nodes]
        if disp is not None:                                  # pressure follows the stretched surface
            xyz = xyz + np.asarray(disp).reshape(nn, 3)[nodes]
        wts, nrm = _face_integrals(xyz)
        f = np.einsum("fgi,fgc,f->fic", wts, nrm, pressure[1:])
        np.add.at(g3, nodes, f)
    lv = np.asarray(loadvertices)
    if len(lv) > 1:
        np.add.at(g3, lv[1:, 0] - 1, np.asarray(vertexloads, dtype=np.float64)[1:])
    lfu = np.asarray(loadfaces_uni)
    if len(lfu) > 1:
        nodes = lfu[1:] - 1
        wts, _ = _face_integrals(nocoord[nodes])
        f = np.einsum("fgi,fc->fic", wts, np.asarray(faceloads, dtype=np.float64)[1:])
        np.add.at(g3, nodes, f)
    le = np.asarray(loadedges)
    if len(le) > 1:
        nodes = le[1:] - 1
        xyz = nocoord[nodes]                                   # (nl, 3, 3)
        el = np.asarray(edgeloads, dtype=np.float64)[1:]
        for xi, w in _GP2:
            shp = np.array([-0.5 * (1.0 - xi) * xi, 0.5 * (1.0 + xi) * xi, (1.0 + xi) * (1.0 - xi)])
            dshp = np.array([xi - 0.5, xi + 0.5, -2.0 * xi])
            xsj = np.linalg.norm(np.einsum("k,lkc->lc", dshp, xyz), axis=1)
            f = shp[None, :, None] * el[:, None, :] * (np.abs(xsj) * w)[:, None, None]
            np.add.at(g3, nodes, f)
    return glv
