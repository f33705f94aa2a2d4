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
    elem_order: Optional[np.ndarray] = None   # set by compact_partition: element k of ``model`` is element
                                              # elem_order[k] of the mesh the caller handed in

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

    def p2p_plan(self, rank: int):
        """Neighbour-only exchange lists of one rank for the peer-memory halo (csrc/fcvm_p2p.cu).

        ``peers``        ranks sharing at least one node with ``rank`` (ascending)
        ``send_ptr/send_node``  per peer: local indices of the nodes shared with it, ascending global id -- the
                         same order on both sides, so position k of the list is the same node for sender and receiver
        ``remote_off``   per peer: where this rank's segment starts inside THAT peer's receive area (in nodes)
        ``n_recv``       nodes of this rank's receive area (segments in ascending peer order)
        ``if_node/if_ptr/if_src``  per local interface node: its contributions in ascending rank order -- -1 for
                         the rank's own value, else the node offset inside the receive area.  Summing in that order
                         gives bit-identical results on every rank that holds the node.
        """
        world = self.world
        g = self.nodes[rank]
        shared = {}                                            # peer -> sorted global ids shared with it
        for q in range(world):
            if q == rank:
                continue
            both = np.intersect1d(g, self.nodes[q], assume_unique=True)
            if both.size:
                shared[q] = both
        peers = sorted(shared)

        def seg_offsets(r):
            off, acc = {}, 0
            gr = self.nodes[r]
            for q in range(world):
                if q == r:
                    continue
                n = np.intersect1d(gr, self.nodes[q], assume_unique=True).size
                if n:
                    off[q] = acc
                    acc += n
            return off, acc

        my_off, n_recv = seg_offsets(rank)
        send_ptr = np.zeros(len(peers) + 1, dtype=np.int32)
        send_node, remote_off = [], []
        for k, q in enumerate(peers):
            loc = np.searchsorted(g, shared[q]).astype(np.int32)
            send_node.append(loc)
            send_ptr[k + 1] = send_ptr[k] + loc.size
            remote_off.append(seg_offsets(q)[0][rank])
        mult = self.multiplicity[g]
        if_node = np.nonzero(mult > 1)[0].astype(np.int32)
        if_ptr = np.zeros(if_node.size + 1, dtype=np.int32)
        # contributions per interface node, ascending rank
        ranks_of = [[] for _ in range(if_node.size)]
        pos_in = {q: np.searchsorted(shared[q], g[if_node]) for q in peers}
        for q in peers:
            here = np.isin(g[if_node], shared[q], assume_unique=True)
            for i in np.nonzero(here)[0]:
                ranks_of[i].append((q, my_off[q] + int(pos_in[q][i])))
        if_src = []
        for i in range(if_node.size):
            ent = sorted(ranks_of[i] + [(rank, -1)])
            if_src.extend(off for _, off in ent)
            if_ptr[i + 1] = len(if_src)
        return dict(peers=np.asarray(peers, dtype=np.int32), send_ptr=send_ptr,
                    send_node=(np.concatenate(send_node) if send_node else np.zeros(0, dtype=np.int32)).astype(np.int32),
                    remote_off=np.asarray(remote_off, dtype=np.int64), n_recv=int(n_recv), if_node=if_node,
                    if_ptr=if_ptr, if_src=np.asarray(if_src, dtype=np.int64))

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
        """Rank-ordered Gauss-point arrays (any number of values per element) -> one array in the element order
        of the mesh the caller handed in."""
        out = np.concatenate([np.asarray(p) for p in parts])
        if self.elem_order is None:
            return out
        ne = self.model.ne
        per = out.reshape(ne, -1)
        back = np.empty_like(per)
        back[self.elem_order] = per
        return back.reshape(out.shape)

    def original_gauss_point(self, gp):
        """Gauss-point numbers 4*element+ip of ``model`` -> numbers in the mesh the caller handed in."""
        gp = np.asarray(gp, dtype=np.int64)
        if self.elem_order is None:
            return gp
        return 4 * self.elem_order[gp // 4] + gp % 4


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


def spatial_order(model: Model, world: int) -> np.ndarray:
    """Element order in which equal contiguous ranges are compact subdomains: recursive coordinate bisection
    of the element centroids (the set is cut across the longest side of its bounding box, into shares
    proportional to the ranks either side; ties go by element number, so the order is reproducible).
    ``order[k]`` is the element that comes k-th.  An unstructured mesh (the reference's Gmsh meshes list their
    elements in the order of the advancing front) otherwise gives every rank a shell that touches all others."""
    cen = model.nocoord[model.elNodes[:, :4] - 1].mean(axis=1)
    order = np.arange(model.ne, dtype=np.int64)

    def cut(lo: int, hi: int, parts: int):
        if parts <= 1 or hi - lo <= 1:
            return
        idx = order[lo:hi]
        c = cen[idx]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        idx = idx[np.lexsort((idx, c[:, axis]))]
        order[lo:hi] = idx
        left = parts // 2
        mid = lo + ((hi - lo) * left) // parts
        cut(lo, mid, left)
        cut(mid, hi, parts - left)

    cut(0, model.ne, max(1, int(world)))
    return order


def compact_partition(model: Model, world: int) -> Partition:
    """``slab_partition`` of the mesh renumbered by ``spatial_order``: for meshes whose element list has no
    spatial order.  Results come back in the caller's element order (``gather_gauss``,
    ``original_gauss_point``); node numbers are untouched.  The one thing that depends on the element order
    itself is which of several EQUAL maxima of csr is reported (np.argmax takes the first, fcVM.py:1546): with a
    renumbered mesh that is the first in the new order."""
    order = spatial_order(model, world)
    ne = model.ne
    # the ranges must be the ones the bisection produced, not equal shares of a rounded count
    bounds = [0]

    def ends(lo, hi, parts):
        if parts <= 1:
            bounds.append(hi)
            return
        left = parts // 2
        mid = lo + ((hi - lo) * left) // parts
        ends(lo, mid, left)
        ends(mid, hi, parts - left)

    ends(0, ne, world)
    renum = dataclasses.replace(model, elNodes=np.ascontiguousarray(model.elNodes[order]),
                                materialbyElement=np.ascontiguousarray(model.materialbyElement[order]))
    part = slab_partition(renum, world, elem_start=np.asarray(bounds, dtype=np.int64))
    part.elem_order = order
    return part


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
            self.p2p = self._attach_p2p(eng)
        _ = _lib

    P2P_SLOT = 16400      # doubles per rank slot of the small exchanges (>= the 16384 coarse unknowns allowed)

    def _attach_p2p(self, eng) -> bool:
        """Peer-memory exchange of the PCG iteration (csrc/fcvm_p2p.cu) between the ranks of one box: arenas
        mapped through CUDA IPC.  Every rank must take the same decision, so failures are agreed on before the
        arenas are used; without them the NCCL path stays.  FCVM_P2P=0 switches it off."""
        import ctypes
        import os

        from ._lib import FcvmError, call, i64p
        i32p = ctypes.POINTER(ctypes.c_int32)
        ok = os.environ.get("FCVM_P2P", "1") != "0" and self.world <= 8
        plan = self.part.p2p_plan(self.rank) if ok else None
        # every arena has the same layout: a sender computes addresses inside its PEER's arena, so the capacity of
        # the receive area (it sets the offset of the second buffer) must not depend on the rank
        halo_cap = max(self.allgather(plan["n_recv"] if ok else 0))
        handle = (ctypes.c_ubyte * 64)()
        if ok:
            try:
                call("fcvm_p2p_create", eng._ctx, halo_cap, self.P2P_SLOT, ctypes.cast(handle, ctypes.c_void_p))
            except FcvmError:
                ok = False
        every = self.allgather((ok, bytes(handle)))
        if not all(e[0] for e in every):
            return False
        blob = (ctypes.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(e[1] for e in every))
        arr = {k: np.ascontiguousarray(plan[k]) for k in ("peers", "send_ptr", "send_node", "remote_off", "if_node",
                                                         "if_ptr", "if_src")}
        try:
            call("fcvm_p2p_attach", eng._ctx, ctypes.cast(blob, ctypes.c_void_p), int(arr["peers"].size),
                 arr["peers"].ctypes.data_as(i32p), arr["send_ptr"].ctypes.data_as(i32p),
                 arr["send_node"].ctypes.data_as(i32p), arr["remote_off"].ctypes.data_as(i64p), int(arr["if_node"].size),
                 arr["if_node"].ctypes.data_as(i32p), arr["if_ptr"].ctypes.data_as(i32p), arr["if_src"].ctypes.data_as(i64p))
            good = True
        except FcvmError as e:
            import warnings
            warnings.warn(f"peer-memory exchange not available, NCCL path kept: {e}")
            good = False
        # the arenas may only be used once EVERY rank has mapped them
        if not all(self.allgather(good)):
            raise FcvmError(-5, "peer-memory arenas mapped on some ranks only")
        return True

    def allgather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def elem_offset(self) -> int:
        return int(self.part.elem_start[self.rank])
