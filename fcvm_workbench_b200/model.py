"""The finite-element input bundle that the reference's ``setUpInput`` returns.

Reference: source code/fcVM.py:122-347 (``setUpInput``).  The field names, the
1-based node numbers, the dummy first row of every load table ("signature for
numba", fcVM.py:260-267) and the CalculiX node order of ``elNodes`` (after the
swap at fcVM.py:338-341) are kept, so the arrays can be handed to either the
reference routines or to this package unchanged.
"""
from __future__ import annotations

import dataclasses
from typing import Dict

import numpy as np


@dataclasses.dataclass
class Model:
    name: str
    elNodes: np.ndarray            # (ne, 10) int64, 1-based, CalculiX order
    nocoord: np.ndarray            # (nn, 3) float64
    fix: Dict[int, float]          # dof -> prescribed value
    fixdof: np.ndarray             # (3nn,) int, 1 = free, 0 = prescribed
    movdof: np.ndarray             # (3nn,) int, 1 = non-zero prescribed value
    materialbyElement: np.ndarray  # (ne, 3) float64: E, nu, density
    noce: np.ndarray               # (nn,) int16: number of elements at a node
    loadfaces: np.ndarray          # (1+nf, 6) int   pressure faces (row 0 is a dummy)
    pressure: np.ndarray           # (1+nf,) float
    loadvertices: np.ndarray       # (1+nv, 1) int
    vertexloads: np.ndarray        # (1+nv, 3) float
    loadedges: np.ndarray          # (1+nl, 3) int
    edgeloads: np.ndarray          # (1+nl, 3) float
    loadfaces_uni: np.ndarray      # (1+nu, 6) int   force-per-area faces
    faceloads: np.ndarray          # (1+nu, 3) float

    @property
    def ne(self) -> int:
        return int(self.elNodes.shape[0])

    @property
    def nn(self) -> int:
        return int(self.nocoord.shape[0])

    def fix_arrays(self):
        """``fix`` as (mask, value) dense arrays over the 3*nn dofs."""
        mask = np.zeros(3 * self.nn, dtype=np.uint8)
        val = np.zeros(3 * self.nn, dtype=np.float64)
        for d, v in self.fix.items():
            mask[d] = 1
            val[d] = v
        return mask, val

    def save_npz(self, path: str) -> None:
        fk = np.fromiter(self.fix.keys(), dtype=np.int64, count=len(self.fix))
        fv = np.fromiter(self.fix.values(), dtype=np.float64, count=len(self.fix))
        np.savez_compressed(
            path, name=np.array(self.name), elNodes=self.elNodes, nocoord=self.nocoord,
            fix_dof=fk, fix_val=fv, fixdof=self.fixdof, movdof=self.movdof,
            materialbyElement=self.materialbyElement, noce=self.noce,
            loadfaces=self.loadfaces, pressure=self.pressure,
            loadvertices=self.loadvertices, vertexloads=self.vertexloads,
            loadedges=self.loadedges, edgeloads=self.edgeloads,
            loadfaces_uni=self.loadfaces_uni, faceloads=self.faceloads)

    @staticmethod
    def load_npz(path: str) -> "Model":
        z = np.load(path, allow_pickle=False)
        fix = {int(d): float(v) for d, v in zip(z["fix_dof"], z["fix_val"])}
        return Model(
            name=str(z["name"]), elNodes=z["elNodes"], nocoord=z["nocoord"], fix=fix,
            fixdof=z["fixdof"], movdof=z["movdof"], materialbyElement=z["materialbyElement"],
            noce=z["noce"], loadfaces=z["loadfaces"], pressure=z["pressure"],
            loadvertices=z["loadvertices"], vertexloads=z["vertexloads"],
            loadedges=z["loadedges"], edgeloads=z["edgeloads"],
            loadfaces_uni=z["loadfaces_uni"], faceloads=z["faceloads"])


def empty_loads():
    """The dummy first rows of the load tables (fcVM.py:260-267)."""
    return dict(
        loadfaces=np.array([[0, 0, 0, 0, 0, 0]]), pressure=np.array([0.0]),
        loadvertices=np.array([[0]]), vertexloads=np.array([[0.0, 0.0, 0.0]]),
        loadedges=np.array([[0, 0, 0]]), edgeloads=np.array([[0.0, 0.0, 0.0]]),
        loadfaces_uni=np.array([[0, 0, 0, 0, 0, 0]]), faceloads=np.array([[0.0, 0.0, 0.0]]))


def finish_bcs(nn: int, dispfaces):
    """fix / fixdof / movdof from a list of (nodes, free-flags, values).

    Same loop as fcVM.py:222-258: a later constraint overwrites an earlier one
    on the same dof, and ``movdof`` marks the non-zero prescribed dofs.
    """
    fix: Dict[int, float] = {}
    fixdof = np.ones(3 * nn, dtype=int)
    movdof = np.zeros(3 * nn, dtype=int)
    for nodes, free, vals in dispfaces:
        for c in range(3):
            if not free[c]:
                for node in nodes:
                    dof = 3 * (int(node) - 1) + c
                    fix[dof] = float(vals[c])
                    fixdof[dof] = 0
    for dof, v in fix.items():
        if v != 0.0:
            movdof[dof] = 1
    return fix, fixdof, movdof


def count_noce(elNodes: np.ndarray, nn: int) -> np.ndarray:
    """Elements per node (fcVM.py:183-185)."""
    return np.bincount(elNodes.ravel() - 1, minlength=nn).astype(np.int16)
