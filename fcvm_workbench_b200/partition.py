"""Element-wise partition of a mesh over the GPUs of one box (one process per GPU).

The Newton path shards by elements: every rank integrates, updates and assembles its own
elements (Gauss-point state never leaves the rank) and holds every node those elements touch.
Nodes touched by more than one rank are *interface* nodes; their nodal sums (internal force,
K*x inside PCG, diagonal blocks, load vector) are completed by one all-reduce over a dense
global interface vector (``fcvm_interface_sum``), and dot products count them once through the
weight ``1 / multiplicity`` (``fcvm_set_interface``).  After an interface sum a shared node
carries bit-identical values on all its ranks.

Local numbering keeps the global order (local nodes sorted by global id, local elements in
global order), so the reference's quirks that depend on numbering -- first maximum of csr,
the last node left out of ``un`` (fcVM.py:1494-1497) -- carry over.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import numpy as np

from .model import Model, empty_loads


@dataclasses.dataclass
class Partition:
    model: Model
    world: int
    elem_start: np.ndarray           # (world+1,) element ranges: rank r owns [elem_start[r], elem_start[r+1])
    nodes: List[np.ndarray]          # per rank: sorted global node ids (0-based) present on the rank
    multiplicity: np.ndarray         # (nn,) number of ranks holding each global node
    if_nodes: np.ndarray             # sorted global ids of the interface nodes (multiplicity > 1)

    @property
    def n_if_global(self) -> int:
        return int(self.if_nodes.size)

    def elements(self, rank: int) -> slice:
        return slice(int(self.elem_start[rank]), int(self.elem_start[rank + 1]))

    def local_model(self, rank: int) -> Model:
        m = self.model
        g = self.nodes[rank]                                   # local -> global (0-based)
        el = m.elNodes[self.elements(rank)]
        loc = np.searchsorted(g, el - 1) + 1                   # 1-based local numbers
        fix = {}
        if m.fix:
            fd = np.fromiter(m.fix.keys(), dtype=np.int64, count=len(m.fix))
            fv = np.fromiter(m.fix.values(), dtype=np.float64, count=len(m.fix))
            nd, cp = fd // 3, fd % 3
            pos = np.searchsorted(g, nd)
            pos[pos >= g.size] = 0
            here = g[pos] == nd
            for d, v in zip(3 * pos[here] + cp[here], fv[here]):
                fix[int(d)] = float(v)
        dof = (3 * g[:, None] + np.arange(3)[None, :]).ravel()
        loads = self._local_loads(rank, g)
        return Model(name=f"{m.name}.part{rank}of{self.world}", elNodes=loc.astype(np.int64),
                     nocoord=np.ascontiguousarray(m.nocoord[g]), fix=fix, fixdof=m.fixdof[dof], movdof=m.movdof[dof],
                     materialbyElement=m.materialbyElement[self.elements(rank)], noce=m.noce[g], **loads)

    # surface loads are integrated by the rank that owns the loaded element (a boundary face belongs to
    # exactly one element); point and line loads by the lowest rank holding their nodes
    def _local_loads(self, rank: int, g: np.ndarray):
        m = self.model
        out = empty_loads()

        def owner_of_nodes(nodes1):
            """lowest rank holding all the given 1-based global nodes of each row"""
            own = np.full(len(nodes1), -1, dtype=np.int64)
            for r in range(self.world - 1, -1, -1):
                has = np.isin(nodes1 - 1, self.nodes[r]).all(axis=1)
                own[has] = r
            return own

        def face_owner(faces):
            own = np.full(len(faces), -1, dtype=np.int64)
            corner_sets = np.sort(faces[:, :3], axis=1)
            tet_faces = ((0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3))
            for r in range(self.world):
                el = m.elNodes[self.elements(r)]
                cand = np.isin(corner_sets, el[:, :4]).all(axis=1) & (own < 0)
                if not cand.any():
                    continue
                keys = set()
                for f in tet_faces:
                    keys.update(map(tuple, np.sort(el[:, list(f)], axis=1)))
                for i in np.nonzero(cand)[0]:
                    if tuple(corner_sets[i]) in keys:
                        own[i] = r
            return own

        def localise(nodes1):
            return (np.searchsorted(g, nodes1 - 1) + 1).astype(nodes1.dtype)

        for tab, val, own_fn in (("loadfaces", "pressure", face_owner), ("loadfaces_uni", "faceloads", face_owner),
                                 ("loadvertices", "vertexloads", owner_of_nodes), ("loadedges", "edgeloads", owner_of_nodes)):
            t, v = getattr(m, tab), getattr(m, val)
            if len(t) <= 1:
                continue
            body = t[1:]
            mine = own_fn(body) == rank
            out[tab] = np.vstack([out[tab], localise(body[mine])]) if mine.any() else out[tab]
            if mine.any():
                vv = v[1:][mine]
                out[val] = np.concatenate([out[val], vv]) if v.ndim == 1 else np.vstack([out[val], vv])
        return out

    # ---- interface maps -------------------------------------------------------------------------------
    def interface(self, rank: int):
        """(dof_weight (3*nn_local), local interface node indices, their global interface slots)"""
        g = self.nodes[rank]
        mult = self.multiplicity[g]
        w = np.repeat(1.0 / mult, 3)
        loc = np.nonzero(mult > 1)[0].astype(np.int64)
        slot = np.searchsorted(self.if_nodes, g[loc]).astype(np.int64)
        return w, loc, slot

    def un_nodes(self, rank: int) -> int:
        """Nodes of this rank that enter ``un``: the reference leaves the last global node out."""
        g = self.nodes[rank]
        return int(g.size - (1 if g[-1] == self.model.nn - 1 else 0))

    # ---- gathering results (host side, rank-ordered lists in) --------------------------------------------
    def gather_nodal(self, parts: List[np.ndarray], ncomp: int = 3) -> np.ndarray:
        out = np.zeros((self.model.nn, ncomp))
        for r, p in enumerate(parts):
            out[self.nodes[r]] = np.asarray(p).reshape(-1, ncomp)
        return out.reshape(-1) if ncomp > 1 else out[:, 0]

    def gather_gauss(self, parts: List[np.ndarray]) -> np.ndarray:
        return np.concatenate([np.asarray(p) for p in parts])


def slab_partition(model: Model, world: int, elem_start: Optional[np.ndarray] = None) -> Partition:
    """Contiguous, equal-sized ranges of the element list.  For the structured meshes of
    ``mesh.box_mesh`` (cells ordered x fastest, z slowest) these are slabs normal to z, so a rank
    shares nodes with at most two neighbours; for an unordered mesh renumber the elements first."""
    ne, nn = model.ne, model.nn
    if world < 1 or world > ne:
        raise ValueError(f"cannot split {ne} elements over {world} ranks")
    if elem_start is None:
        elem_start = (np.arange(world + 1, dtype=np.int64) * ne) // world
    elem_start = np.asarray(elem_start, dtype=np.int64)
    nodes, mult = [], np.zeros(nn, dtype=np.int32)
    for r in range(world):
        g = np.unique(model.elNodes[int(elem_start[r]):int(elem_start[r + 1])]) - 1
        nodes.append(g.astype(np.int64))
        mult[g] += 1
    if (mult == 0).any():
        # nodes no element refers to (the reference tolerates them): give them to rank 0
        orphan = np.nonzero(mult == 0)[0]
        nodes[0] = np.union1d(nodes[0], orphan)
        mult[orphan] = 1
    return Partition(model=model, world=world, elem_start=elem_start, nodes=nodes, multiplicity=mult,
                     if_nodes=np.nonzero(mult > 1)[0].astype(np.int64))


class Comm:
    """Communicator of one rank: binds the partition's interface maps and an NCCL communicator to
    an ``fcVM.Engine``.  The NCCL unique id travels through ``torch.distributed`` (any backend);
    small host-side reductions (argmax of csr per load step, result gathering) use it too."""

    def __init__(self, part: Partition, rank: int, world: int, dist=None):
        self.part, self.rank, self.world = part, rank, world
        if dist is None:
            import torch.distributed as dist
        self.dist = dist

    def attach(self, eng):
        import ctypes

        from . import _lib
        from ._lib import call, f64p, i64p
        w, loc, slot = self.part.interface(self.rank)
        w = np.ascontiguousarray(w)
        call("fcvm_set_interface", eng._ctx, w.ctypes.data_as(f64p), int(loc.size), loc.ctypes.data_as(i64p),
             slot.ctypes.data_as(i64p), self.part.n_if_global)
        call("fcvm_set_un_nodes", eng._ctx, self.part.un_nodes(self.rank))
        if self.world > 1:
            buf = (ctypes.c_ubyte * 128)()
            if self.rank == 0:
                call("fcvm_comm_unique_id", ctypes.cast(buf, ctypes.c_void_p))
            box = [bytes(buf)]
            self.dist.broadcast_object_list(box, src=0)
            idb = (ctypes.c_ubyte * 128).from_buffer_copy(box[0])
            call("fcvm_comm_init", eng._ctx, ctypes.cast(idb, ctypes.c_void_p), self.rank, self.world)
        _ = _lib

    def allgather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def elem_offset(self) -> int:
        return int(self.part.elem_start[self.rank])
