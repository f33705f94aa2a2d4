"""Result files of an analysis: the ``.out`` table and the ``.vtk`` grid the workbench writes.

``write_out`` reproduces the text file of ``fcVM.FCMacro:212-262`` character for character
(same header lines, same column formats); ``write_vtk`` writes the unstructured grid of
``exportVTK`` (fcVM.py:2903-2950) as a legacy VTK 5.1 binary file with the same point-data arrays
under the names meshio gives them (blanks -> underscores), so ParaView state files made for the
reference's output open unchanged.  The reference goes through pyvista + meshio; neither is
needed here.
"""
from __future__ import annotations

import numpy as np

_RULE = "\n" + "=" * 121 + "\n\n"


def write_out(path, name, ne, nn, gnl, nstep, loadsum, res, x=None, eigenval=None):
    """``res``: the dictionary ``fcVM.calcDisp`` returns; ``x``: Gauss-point coordinates (4*ne, 3) or
    None to take ``res['x_crip']`` (coordinates of the listed points only)."""
    crip = np.asarray(res["crip"], dtype=np.int64)
    xc = np.asarray(x)[crip] if x is not None else np.asarray(res["x_crip"])
    with open(path, "w") as f:
        f.write("model name:{0: >50}\n".format(name))
        f.write("No. of elements:{0: >45}\n".format(ne))
        f.write("No. of Degrees of freedom:{0: >35}\n".format(nn))
        if gnl == "GNLY":
            kind = "elastic buckling analysis" if float(nstep) == 1.0 else "elastic-plastic, geometric non-linear"
            f.write("analysis type:{0: >47}\n".format(kind))
            f.write("elastic buckling factors:{0: >36}\n".format(str(eigenval)))
        elif float(nstep) == 1.0:
            f.write("analysis type: elastic\n")
        else:
            f.write("analysis type: elastic-plastic, geometric linear\n")
        f.write(_RULE)
        for axis, v in zip("xyz", loadsum):
            f.write("Sum of loads {0}-direction: {1: >15.2e}\n".format(axis, v))
        f.write(_RULE)
        f.write("{0: >8}{1: >10}{2: >10}{3: >10}{4: >10}{5: >10}{6: >10}{7: >10}{8: >10}{9: >10}{10: >10}{11: >10}\n".format(
            "Gauss point", "x", "y", "z", "load", "disp", "peeq", "pressure", "svmises", "triax", "eps_cr", "csr_max"))
        for i in range(len(crip)):
            f.write("{0: 11d}{1: >10.2e}{2: >10.2e}{3: >10.2e}{4: >10.2e}{5: >10.2e}{6: >10.2e}{7: >10.2e}{8: >10.2e}"
                    "{9: >10.2e}{10: >10.2e}{11: >10.2e}\n".format(
                        int(crip[i]), xc[i][0], xc[i][1], xc[i][2], res["lout"][i], res["un"][i], res["peeqplot"][i],
                        res["pplot"][i], res["svmplot"][i], res["triaxplot"][i], res["ecrplot"][i], res["csrplot"][i]))
        f.write(_RULE)


def principal_stresses(tet10stress):
    """calculate_principal_stress (fcVM.py:2953-2994), vectorised: values descending, vectors scaled by them."""
    s = np.asarray(tet10stress, dtype=np.float64).reshape(-1, 6)
    sig = np.empty((len(s), 3, 3))
    sig[:, 0, 0], sig[:, 1, 1], sig[:, 2, 2] = s[:, 0], s[:, 1], s[:, 2]
    sig[:, 0, 1] = sig[:, 1, 0] = s[:, 3]
    sig[:, 0, 2] = sig[:, 2, 0] = s[:, 4]
    sig[:, 1, 2] = sig[:, 2, 1] = s[:, 5]
    w, v = np.linalg.eigh(sig)                     # ascending
    w, v = w[:, ::-1], v[:, :, ::-1]
    vec = [w[:, k, None] * v[:, :, k] for k in range(3)]
    return w[:, 0].copy(), w[:, 1].copy(), w[:, 2].copy(), vec[0], vec[1], vec[2]


def _be(a, dt):
    return np.ascontiguousarray(a, dtype=np.dtype(dt).newbyteorder(">")).tobytes()


def write_vtk(path, elNodes, nocoord, dis, tet10stress, tet10peeq, tet10csr, tet10svm, tet10triax):
    """Quadratic tetrahedra (VTK cell type 24) with the point data of ``exportVTK``."""
    el = np.asarray(elNodes, dtype=np.int64) - 1
    xyz = np.asarray(nocoord, dtype=np.float64)
    ne, nn = len(el), len(xyz)
    s1, s2, s3, v1, v2, v3 = principal_stresses(tet10stress)
    scal = [("Critical_Strain_Ratio\n", tet10csr), ("Equivalent_Plastic_Strain\n", tet10peeq),
            ("von_Mises_Stress\n", tet10svm), ("Triaxiality\n", tet10triax)]
    fields = ([(n, np.asarray(a).reshape(nn, 1)) for n, a in scal]
              + [("Displacement", np.asarray(dis).reshape(nn, 3)), ("Stress_Tensor", np.asarray(tet10stress).reshape(nn, 6)),
                 ("Major_Principal_Stress\n", s1.reshape(nn, 1)), ("Intermediate_Principal_Stress\n", s2.reshape(nn, 1)),
                 ("Minor_Principal_Stress\n", s3.reshape(nn, 1)), ("Major_Principal_Stress_Vector", v1),
                 ("Intermediate_Principal_Stress_Vector", v2), ("Minor_Principal_Stress_Vector", v3)])
    with open(path, "wb") as f:
        w = lambda t: f.write(t.encode("ascii"))
        w("# vtk DataFile Version 5.1\nwritten by fcvm_workbench_b200\nBINARY\nDATASET UNSTRUCTURED_GRID\n")
        w(f"POINTS {nn} double\n")
        f.write(_be(xyz, "f8"))
        w(f"\nCELLS {ne + 1} {10 * ne}\nOFFSETS vtktypeint64\n")
        f.write(_be(np.arange(ne + 1, dtype=np.int64) * 10, "i8"))
        w("\nCONNECTIVITY vtktypeint64\n")
        f.write(_be(el, "i8"))
        w(f"\nCELL_TYPES {ne}\n")
        f.write(_be(np.full(ne, 24), "i4"))
        w(f"\nPOINT_DATA {nn}\nFIELD FieldData {len(fields)}\n")
        for name, a in fields:
            w(f"{name} {a.shape[1]} {nn} double\n")
            f.write(_be(a, "f8"))
            w("\n")


def read_vtk_points_and_fields(path):
    """Minimal reader of the files ``write_vtk`` (and meshio, for the reference's output) writes;
    returns (points, cells, {field name: array}).  Used by the tests."""
    data = open(path, "rb").read()
    pos = 0

    def line():
        nonlocal pos
        end = data.index(b"\n", pos)
        s = data[pos:end].decode("ascii", "replace")
        pos = end + 1
        return s

    def block(n, dt):
        nonlocal pos
        a = np.frombuffer(data, dtype=np.dtype(dt).newbyteorder(">"), count=n, offset=pos)
        pos += a.nbytes
        return a.astype(np.dtype(dt))

    for _ in range(4):
        line()
    nn = int(line().split()[1])
    pts = block(3 * nn, "f8").reshape(nn, 3)
    s = line()
    while not s.startswith("CELLS"):
        s = line()
    noff = int(s.split()[1])
    line()
    off = block(noff, "i8")
    s = line()
    while not s.startswith("CONNECTIVITY"):
        s = line()
    conn = block(int(off[-1]), "i8").reshape(noff - 1, -1)
    s = line()
    while not s.startswith("CELL_TYPES"):
        s = line()
    block(int(s.split()[1]), "i4")
    s = line()
    while not s.startswith("FIELD"):
        s = line()
    out = {}
    for _ in range(int(s.split()[2])):
        s = line()
        while s.strip() == "":
            s = line()
        parts = s.split()
        if len(parts) < 4:                     # meshio keeps the "\n" of the reference's array names
            name = parts[0]
            ncomp, nt, _ = line().split()
        else:
            name, ncomp, nt = parts[0], parts[1], parts[2]
        out[name] = block(int(ncomp) * int(nt), "f8").reshape(int(nt), int(ncomp))
    return pts, conn, out
