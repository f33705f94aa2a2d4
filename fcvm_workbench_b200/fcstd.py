"""Headless reader for the reference's model files (``freeCAD files/*.FCStd``).

A ``.FCStd`` document is a zip archive: ``Document.xml`` holds the analysis
objects (constraints, material), ``FemMesh.unv`` the second-order Gmsh/Netgen
mesh in I-DEAS universal format.  The reference pulls the same information out
of a live FreeCAD session (source code/fcVM.py:122-347, ``setUpInput``); here it
is read straight from the archive so that an analysis can be driven without
FreeCAD.  Geometry references ("Box.Face6") are resolved for the shapes whose
faces are known in closed form (``Part::Box``) and, for other planar-faced
solids, from the point lists FreeCAD stores with each constraint.

Node orders
-----------
UNV writes second-order cells interlaced (corner, mid, corner, ...).  SMESH --
and therefore ``FemMesh.getElementNodes`` which the reference calls
(fcVM.py:162-164) -- lists corners first.  ``_UNV_TO_SMESH`` undoes the
interlacing; ``setUpInput`` then swaps (1,2), (4,6), (8,9) to reach the
CalculiX order used by the element routines (fcVM.py:338-341).
"""
from __future__ import annotations

import re
import xml.etree.ElementTree as ET
import zipfile
from typing import Dict, List, Tuple

import numpy as np

from .model import Model, count_noce, empty_loads, finish_bcs

# unv[k] = smesh[_INTERLACE[k]]
_INTERLACE = {
    118: (0, 4, 1, 5, 2, 6, 7, 8, 9, 3),   # parabolic tetrahedron
    42: (0, 3, 1, 4, 2, 5),                # parabolic triangle
    22: (0, 2, 1), 24: (0, 2, 1),          # parabolic beam
}
_BEAM_TYPES = (11, 21, 22, 24)


def read_unv(text: str):
    """Nodes and second-order cells of a UNV mesh.

    Returns ``(node_ids, coords, cells)`` where ``cells[type]`` is a list of
    ``(element_id, nodes-in-SMESH-order)``.
    """
    blocks = re.split(r"^\s{4}-1\s*$", text, flags=re.M)
    node_ids: List[int] = []
    coords: List[Tuple[float, float, float]] = []
    cells: Dict[int, list] = {}
    for blk in blocks:
        lines = [ln for ln in blk.strip("\n").split("\n") if ln.strip() != ""]
        if not lines:
            continue
        ds = lines[0].strip()
        if ds == "2411":
            for i in range(1, len(lines) - 1, 2):
                node_ids.append(int(lines[i].split()[0]))
                x, y, z = (float(v.replace("D", "E")) for v in lines[i + 1].split())
                coords.append((x, y, z))
        elif ds == "2412":
            i = 1
            while i < len(lines):
                head = lines[i].split()
                eid, etype, nnod = int(head[0]), int(head[1]), int(head[5])
                i += 1
                if etype in _BEAM_TYPES:
                    i += 1                      # beam orientation record
                nodes: List[int] = []
                while len(nodes) < nnod:
                    nodes.extend(int(v) for v in lines[i].split())
                    i += 1
                if etype in _INTERLACE:
                    il = _INTERLACE[etype]
                    smesh = [0] * nnod
                    for k, n in enumerate(nodes):
                        smesh[il[k]] = n
                    nodes = smesh
                cells.setdefault(etype, []).append((eid, nodes))
    return np.asarray(node_ids, dtype=np.int64), np.asarray(coords, dtype=np.float64), cells


def _prop(obj, name):
    for p in obj.find("Properties").findall("Property"):
        if p.get("name") == name:
            return p
    return None


def _float(obj, name, default=0.0):
    p = _prop(obj, name)
    if p is None:
        return default
    f = p.find("Float")
    return float(f.get("value")) if f is not None else default


def _bool(obj, name, default=False):
    p = _prop(obj, name)
    if p is None:
        return default
    b = p.find("Bool")
    return (b.get("value") == "true") if b is not None else default


def _refs(obj):
    p = _prop(obj, "References")
    out = []
    if p is not None:
        for ln in p.iter("Link"):
            out.append((ln.get("obj"), ln.get("sub")))
    return out


def _quantity(s: str) -> Tuple[float, str]:
    m = re.match(r"\s*([-+0-9.eE]+)\s*(.*)", s)
    return float(m.group(1)), m.group(2).strip()


def _to_mpa(s: str) -> float:
    v, u = _quantity(s)
    return v * {"MPa": 1.0, "GPa": 1.0e3, "kPa": 1.0e-3, "Pa": 1.0e-6, "N/mm^2": 1.0, "": 1.0}[u]


def _to_kg_mm3(s: str) -> float:
    v, u = _quantity(s)
    return v * {"kg/m^3": 1.0e-9, "kg/mm^3": 1.0, "g/cm^3": 1.0e-6, "kg/dm^3": 1.0e-6, "": 1.0}[u]


class _BoxShape:
    """Faces of a ``Part::Box`` in FreeCAD's numbering (Face1..Face6 =
    x-min, x-max, y-min, y-max, z-min, z-max), placed at the origin."""

    def __init__(self, length, width, height):
        self.ext = (length, width, height)

    def face(self, sub: str):
        k = int(sub[4:]) - 1
        axis, side = divmod(k, 2)
        return axis, (self.ext[axis] if side else 0.0)

    def face_area(self, sub: str) -> float:
        axis, _ = self.face(sub)
        a = [self.ext[i] for i in range(3) if i != axis]
        return a[0] * a[1]

    def vertex(self, sub: str):
        """Vertex1..Vertex8 of a box: x from bit 2, y from bit 1 of (number - 1), z = height when bit 0 is
        clear (Vertex1 = (0, 0, H), Vertex2 = (0, 0, 0), Vertex4 = (0, W, 0), Vertex6 = (L, 0, 0) -- the
        coordinates FreeCAD stores with the constraints of VM_Uniaxial_Tension_Example.FCStd)."""
        k = int(sub[6:]) - 1
        return np.array([self.ext[0] if k & 4 else 0.0, self.ext[1] if k & 2 else 0.0, 0.0 if k & 1 else self.ext[2]])

    def nodes_on(self, sub: str, coords: np.ndarray, tol: float) -> np.ndarray:
        if sub.startswith("Vertex"):
            return np.nonzero(np.linalg.norm(coords - self.vertex(sub), axis=1) <= tol)[0]
        if not sub.startswith("Face"):
            raise NotImplementedError(f"'{sub}': only faces and vertices of a Part::Box are resolved headlessly")
        axis, val = self.face(sub)
        return np.nonzero(np.abs(coords[:, axis] - val) <= tol)[0]


class _SampledPlanarFaces:
    """Faces of an arbitrary solid, as far as a constraint needs them: FreeCAD stores with every constraint a
    sample of points (and normals) on the faces it references (properties ``Points`` / ``Normals``).  For planar
    faces that is enough to find the mesh nodes on them without a CAD kernel: the samples are grouped by plane,
    and the nodes of a face are the mesh nodes in that plane within the bounding box of its samples (the samples
    include the face boundary; the body meets the plane only in the face)."""

    def __init__(self, points: np.ndarray, normals: np.ndarray):
        self.planes = []
        n = normals / np.maximum(np.linalg.norm(normals, axis=1, keepdims=True), 1e-300)
        d = np.einsum("ij,ij->i", points, n)
        scale = max(float(np.abs(points).max()), 1.0)
        key = np.round(np.c_[n, d / scale], 6)
        for k in np.unique(key, axis=0):
            sel = np.all(key == k, axis=1)
            pts = points[sel]
            # a planar face contributes a whole grid of samples with one normal; a curved face scatters its
            # samples over as many (normal, offset) pairs as it has points
            if len(pts) < 3 or np.linalg.matrix_rank(pts - pts[0], tol=1e-9 * scale) > 2:
                raise NotImplementedError("constraint on a curved face: only planar faces are resolved headlessly")
            self.planes.append((n[sel][0], float(d[sel].mean()), pts.min(axis=0), pts.max(axis=0)))

    def nodes(self, coords: np.ndarray, tol: float) -> np.ndarray:
        on = np.zeros(len(coords), dtype=bool)
        for nrm, d, lo, hi in self.planes:
            on |= (np.abs(coords @ nrm - d) <= tol) & np.all((coords >= lo - tol) & (coords <= hi + tol), axis=1)
        return np.nonzero(on)[0]


def _vector_list(z, obj, name):
    """A ``PropertyVectorList`` stored as a separate binary file of the archive (uint32 count, 3 doubles each)."""
    p = _prop(obj, name)
    node = p.find("VectorList") if p is not None else None
    if node is None or not node.get("file"):
        return None
    raw = z.read(node.get("file"))
    n = int(np.frombuffer(raw[:4], dtype="<u4")[0])
    dt = "<f8" if len(raw) == 4 + 24 * n else "<f4"
    return np.frombuffer(raw[4:], dtype=dt, count=3 * n).reshape(n, 3).astype(np.float64)


def read_fcstd(path: str) -> Model:
    """Build the ``setUpInput`` arrays from a ``.FCStd`` archive."""
    with zipfile.ZipFile(path) as z:
        doc = ET.fromstring(z.read("Document.xml"))
        unv = z.read("FemMesh.unv").decode()
        sampled = {}
        for o in doc.find("ObjectData").findall("Object"):
            pts, nrm = _vector_list(z, o, "Points"), _vector_list(z, o, "Normals")
            if pts is not None and nrm is not None and len(pts) == len(nrm) and len(pts) > 0:
                sampled[o.get("name")] = (pts, nrm)
    label = None
    props = doc.find("Properties")
    if props is not None:
        for p in props.findall("Property"):
            if p.get("name") == "Label":
                label = p.find("String").get("value")
    node_ids, nocoord, cells = read_unv(unv)
    if len(node_ids) == 0 or 118 not in cells:
        raise ValueError(f"{path}: the document holds no second-order tetrahedral mesh")
    if not np.array_equal(node_ids, np.arange(1, len(node_ids) + 1)):
        raise ValueError("node labels must be 1..nn (the reference indexes nocoord[node-1])")
    nn = len(node_ids)

    objects = {o.get("name"): o for o in doc.find("ObjectData").findall("Object")}
    types = {o.get("name"): o.get("type") for o in doc.find("Objects").findall("Object")}

    shapes = {}
    for name, typ in types.items():
        if typ == "Part::Box":
            o = objects[name]
            shapes[name] = _BoxShape(_float(o, "Length"), _float(o, "Width"), _float(o, "Height"))

    # volume elements, ascending element id (FemMesh.Volumes), then the reference's swap
    vol = sorted(cells[118], key=lambda t: t[0])
    elNodes = np.asarray([n for _, n in vol], dtype=np.int64)
    elNodes[:, [1, 2]] = elNodes[:, [2, 1]]
    elNodes[:, [4, 6]] = elNodes[:, [6, 4]]
    elNodes[:, [8, 9]] = elNodes[:, [9, 8]]
    ne = len(elNodes)

    E, nu, rho = 210000.0, 0.3, 7.9e-6
    for name, typ in types.items():
        if typ.startswith("App::MaterialObject"):
            m = {it.get("key"): it.get("value") for it in _prop(objects[name], "Material").iter("Item")}
            E = _to_mpa(m["YoungsModulus"])
            nu = float(m["PoissonRatio"])
            rho = _to_kg_mm3(m.get("Density", "0 kg/m^3"))
    materialbyElement = np.tile(np.array([E, nu, rho]), (ne, 1))

    span = float(np.max(nocoord.max(axis=0) - nocoord.min(axis=0)))
    tol = 1.0e-6 * span

    def shape_of(obj_name):
        if obj_name not in shapes:
            raise NotImplementedError(
                f"geometry reference to '{obj_name}' ({types.get(obj_name)}): only Part::Box faces are "
                "resolved headlessly; export the model with fcvm_workbench_b200.model.Model.save_npz instead")
        return shapes[obj_name]

    tri = sorted(cells.get(42, []), key=lambda t: t[0])

    def faces_on(obj_name, sub):
        on = set((shape_of(obj_name).nodes_on(sub, nocoord, tol) + 1).tolist())
        return [n for _, n in tri if all(k in on for k in n)]

    dispfaces = []
    loads = empty_loads()
    lf, pr = loads["loadfaces"].tolist(), loads["pressure"].tolist()
    lfu, fl = loads["loadfaces_uni"].tolist(), loads["faceloads"].tolist()
    for name in [o.get("name") for o in doc.find("Objects").findall("Object")]:
        typ, o = types[name], objects[name]
        if typ in ("Fem::ConstraintFixed", "Fem::ConstraintDisplacement"):
            if typ == "Fem::ConstraintFixed":
                free, vals = [False] * 3, [0.0] * 3
            else:
                free = [_bool(o, "xFree", True), _bool(o, "yFree", True), _bool(o, "zFree", True)]
                vals = [_float(o, "xDisplacement"), _float(o, "yDisplacement"), _float(o, "zDisplacement")]
            bc = []
            refs = _refs(o)
            if refs and any(on not in shapes for on, _ in refs) and name in sampled:
                # not a Part::Box: vertices are the stored points themselves, planar faces are recovered from
                # the samples stored with the constraint
                if all(sub.startswith("Vertex") for _, sub in refs):
                    for pnt in sampled[name][0]:
                        bc.extend((np.nonzero(np.linalg.norm(nocoord - pnt, axis=1) <= tol)[0] + 1).tolist())
                elif all(sub.startswith("Face") for _, sub in refs):
                    bc.extend((_SampledPlanarFaces(*sampled[name]).nodes(nocoord, tol) + 1).tolist())
                else:
                    raise NotImplementedError("mixed vertex / edge / face references on a shape that is not a Part::Box")
                refs = []
            for obj_name, sub in refs:
                bc.extend((shape_of(obj_name).nodes_on(sub, nocoord, tol) + 1).tolist())
            bc = list(dict.fromkeys(bc))
            if bc:
                dispfaces.append((bc, free, vals))
        elif typ == "Fem::ConstraintPressure":
            sign = 1 if _bool(o, "Reversed") else -1
            p = _float(o, "Pressure")
            for obj_name, sub in _refs(o):
                for nodes in faces_on(obj_name, sub):
                    lf.append(nodes)
                    pr.append(sign * p)
        elif typ == "Fem::ConstraintForce":
            F = _float(o, "Force")
            dv = _prop(o, "DirectionVector").find("PropertyVector")
            d = [float(dv.get("valueX")), float(dv.get("valueY")), float(dv.get("valueZ"))]
            refs = _refs(o)
            A = sum(shape_of(on).face_area(sub) for on, sub in refs)
            for obj_name, sub in refs:
                for nodes in faces_on(obj_name, sub):
                    lfu.append(nodes)
                    fl.append([F * d[0] / A, F * d[1] / A, F * d[2] / A])

    fix, fixdof, movdof = finish_bcs(nn, dispfaces)
    loads["loadfaces"], loads["pressure"] = np.array(lf), np.array(pr)
    loads["loadfaces_uni"], loads["faceloads"] = np.array(lfu), np.array(fl)
    return Model(name=label or "model", elNodes=elNodes, nocoord=nocoord, fix=fix, fixdof=fixdof,
                 movdof=movdof, materialbyElement=materialbyElement, noce=count_noce(elNodes, nn), **loads)
