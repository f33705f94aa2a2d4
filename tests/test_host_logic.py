"""Host-side logic that needs no GPU: control files, model readers, meshes, load vectors."""
import os

import numpy as np
import pytest

from fcvm_workbench_b200.control import Control, read_control
from fcvm_workbench_b200.loads import surface_load_vector
from fcvm_workbench_b200.mesh import box_mesh, cube_model
from fcvm_workbench_b200.model import Model

from _golden import load, model_of


def test_control_roundtrip(tmp_path):
    c = Control(sig_yield=355.0, nstep=7, gnl="GNLY", disp_output="incremental")
    p = tmp_path / "x.inp"
    c.write(str(p))
    assert read_control(str(p)) == c


def test_control_short_file_gets_reference_defaults(tmp_path):
    p = tmp_path / "old.inp"
    p.write_text("\n".join(["100", "0.0", "0.0", "0.0", "25", "20", "0.001", "1.2", "2.0", "1.2", "1.2",
                            "incremental", "0.25", "0.0", "3", "PEEQ", "unaveraged"]) + "\n\n")
    c = read_control(str(p))
    assert c.nstep == 25 and c.gnl == "GNLN" and c.target_LF == 3.0


def test_control_missing_field_is_an_error(tmp_path):
    p = tmp_path / "bad.inp"
    p.write_text("100\n0.0\n")
    with pytest.raises(ValueError):
        read_control(str(p))


@pytest.mark.parametrize("n", [(1, 1, 1), (3, 2, 4)])
def test_box_mesh_is_conforming_and_positive(n):
    el, xyz = box_mesh(*n, 3.0, 2.0, 4.0)
    assert el.shape == (6 * n[0] * n[1] * n[2], 10)
    assert xyz.shape[0] == (2 * n[0] + 1) * (2 * n[1] + 1) * (2 * n[2] + 1)
    assert len(np.unique(el)) == xyz.shape[0] and el.min() == 1
    X = xyz[el - 1]
    pairs = {4: (0, 1), 5: (1, 2), 6: (0, 2), 7: (0, 3), 8: (1, 3), 9: (2, 3)}
    for k, (a, b) in pairs.items():
        assert np.allclose(X[:, k], 0.5 * (X[:, a] + X[:, b]))
    vol = np.einsum("ei,ei->e", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0]) / 6
    assert (vol > 0).all() and vol.sum() == pytest.approx(24.0)


def test_model_npz_roundtrip(tmp_path):
    m = cube_model(2, mode="force", top_disp=10.0)
    p = str(tmp_path / "m.npz")
    m.save_npz(p)
    r = Model.load_npz(p)
    assert np.array_equal(r.elNodes, m.elNodes) and r.fix == m.fix and np.array_equal(r.loadfaces_uni, m.loadfaces_uni)


def test_surface_loads_match_reference_glv():
    """glv of the reference's calcGSM for the tensile model (Face6 force) and the force cube."""
    for name in ("tensile", "cube2_force"):
        z = load(name)
        m = model_of(z)
        if float(z["c_grav_z"]) != 0.0:
            continue                       # gravity part is integrated on the device
        glv = surface_load_vector(m.nocoord, m.loadfaces, m.pressure, m.loadvertices, m.vertexloads, m.loadedges,
                                  m.edgeloads, m.loadfaces_uni, m.faceloads)
        assert np.abs(glv - z["r_glv"]).max() < 1e-9 * np.abs(z["r_glv"]).max()


def test_surface_loads_all_kinds_vs_oracle(oracle):
    m = cube_model(2, mode="force", top_disp=300.0)
    rng = np.random.default_rng(1)
    xyz = m.nocoord + rng.uniform(-.05, .05, m.nocoord.shape)
    lf = np.vstack([m.loadfaces, m.loadfaces_uni[1:4]])
    pr = np.array([0., 3., -2., 5.])
    le = np.array([[0, 0, 0], [1, 3, 2], [3, 5, 4]])
    el = np.array([[0, 0, 0], [1., 2, 3], [4, 5, 6.]])
    lv = np.array([[0], [7], [9]])
    vl = np.array([[0, 0, 0], [1., 0, 2], [0, 3, 1.]])
    for disp in (None, rng.normal(0, .01, 3 * m.nn)):
        a = surface_load_vector(xyz, lf, pr, lv, vl, le, el, m.loadfaces_uni, m.faceloads, disp=disp)
        b = oracle.load_vector(xyz, lf, pr, lv, vl, le, el, m.loadfaces_uni, m.faceloads, disp=disp)
        assert np.abs(a - b).max() < 1e-12 * np.abs(b).max()


def test_fcstd_reader_on_reference_model():
    path = "/root/reference/freeCAD files/tensile.FCStd"
    if not os.path.isfile(path):
        pytest.skip("reference models not present on this machine")
    from fcvm_workbench_b200.fcstd import read_fcstd
    m = read_fcstd(path)
    z = load("tensile")
    assert np.array_equal(m.elNodes, z["m_elNodes"]) and np.allclose(m.nocoord, z["m_nocoord"])
    assert sorted(m.fix) == sorted(int(d) for d in z["m_fix_dof"])


def test_plate_with_hole_mesh_is_valid_and_curved():
    from fcvm_workbench_b200.mesh import plate_with_hole_model
    m = plate_with_hole_model(3, 6, 1)
    xyz = m.nocoord[m.elNodes - 1]                                          # (ne, 10, 3)
    v = np.einsum("ei,ei->e", np.cross(xyz[:, 1] - xyz[:, 0], xyz[:, 2] - xyz[:, 0]), xyz[:, 3] - xyz[:, 0])
    assert (v > 0).all()                                                    # positively oriented corners
    mid01 = 0.5 * (xyz[:, 0] + xyz[:, 1])
    assert np.abs(xyz[:, 4] - mid01).max() > 1e-3                           # mid-side nodes follow the curved map
    r = np.hypot(m.nocoord[:, 0], m.nocoord[:, 1])
    assert abs(r.min() - 10.0) < 1e-9 and m.nocoord[:, 0].max() == pytest.approx(50.0)
    assert m.movdof.sum() > 0 and len(m.fix) > 0


def test_deflation_box_grid():
    from fcvm_workbench_b200.fcVM import deflation_boxes
    # 1M-element cube of the bench: 55 cells of 10/55 per edge -> 10 x 10 x 10 boxes for 6144 unknowns
    g, h = deflation_boxes([0, 0, 0], [10, 10, 10], [10 / 55] * 3, 6144)
    assert tuple(g) == (10, 10, 10) and np.allclose(h, 1.0)
    # a slab: boxes stay near-cubic
    g, h = deflation_boxes([0, 0, 0], [10, 10, 80], [0.2, 0.2, 0.2], 6 * 512)
    assert g[2] > 4 * g[0] and abs(np.prod(g) - 512) < 200
    # coarse meshes: never narrower than two elements, so tiny meshes end with a single box
    g, _ = deflation_boxes([0, 0, 0], [4, 4, 4], [2, 2, 2], 6144)
    assert tuple(g) == (1, 1, 1)
    g, h = deflation_boxes([0, 0, 0], [6, 6, 6], [1, 1, 1], 10 ** 6)
    assert tuple(g) == (2, 2, 2) and (h >= 2.0).all()
    # the dense coarse problem is capped at 16384 unknowns
    g, _ = deflation_boxes([0, 0, 0], [1, 1, 1], [1e-3] * 3, 10 ** 6)
    assert 6 * np.prod(g) <= 16384
    # explicit grid is clipped by the same rules
    g, _ = deflation_boxes([0, 0, 0], [10, 10, 10], [1, 1, 1], grid=(8, 8, 8))
    assert tuple(g) == (4, 4, 4)


def test_host_engine_vector_algebra_is_in_place_and_exact():
    """hostpath.HostEngine's numpy algebra (the reference's vector algebra between the heavy calls)."""
    from fcvm_workbench_b200 import hostpath
    H = hostpath.HostEngine.__new__(hostpath.HostEngine)          # no GPU: only the host-side methods
    H._w, H.comm = None, None
    rng = np.random.default_rng(0)
    H._nodal = {hostpath._fc.FIXDOF: (rng.random(1000) > 0.2).astype(float)}
    for a, b in ((2.0, 0.0), (0.0, 0.5), (1.5, 1.0), (1.5, -0.3)):
        x, y = rng.normal(size=1000), rng.normal(size=1000)
        ref = a * x + b * y
        H.axpby(a, x, b, y)
        assert np.allclose(y, ref, rtol=0, atol=1e-15)
    for a, b, c in ((1.0, 0.7, 1.0), (2.0, 0.7, 0.0), (0.25, -0.25, 0.0), (1.5, 2.5, 0.5)):
        x, y, z = (rng.normal(size=1000) for _ in range(3))
        ref = a * x + b * y + c * z
        H.axpbypcz(a, x, b, y, c, z)
        assert np.allclose(z, ref, rtol=0, atol=1e-14)
    glv, qin, r = (rng.normal(size=1000) for _ in range(3))
    ref = H._nodal[hostpath._fc.FIXDOF] * (1.3 * glv - qin)
    n = H.residual(1.3, glv, qin, r)
    assert np.allclose(r, ref, rtol=0, atol=1e-15) and abs(n - np.linalg.norm(ref)) < 1e-12
    d = rng.normal(size=3 * 7)
    H.ndof, H._un_nodes = 21, None
    assert H.max_node_disp(d) == pytest.approx(np.sqrt((d[:18].reshape(-1, 3) ** 2).sum(axis=1).max()))   # last node left out


def test_fcstd_reader_resolves_box_vertices_of_the_uniaxial_example():
    """BASELINE config 0: constraints on Vertex2 / Vertex4 / Vertex6 of a Part::Box and pressures on all faces."""
    path = "/root/reference/freeCAD files/VM_Uniaxial_Tension_Example.FCStd"
    if not os.path.isfile(path):
        pytest.skip("reference models not present on this machine")
    from fcvm_workbench_b200.fcstd import read_fcstd, _BoxShape
    m = read_fcstd(path)
    z = load("vm_uniaxial_tension")
    assert np.array_equal(m.elNodes, z["m_elNodes"]) and np.allclose(m.nocoord, z["m_nocoord"])
    # statically determinate support: origin fully fixed, (0,10,0) in x and z, (10,0,0) in z
    at = lambda p: int(np.nonzero(np.linalg.norm(m.nocoord - np.array(p), axis=1) < 1e-9)[0][0])
    o, a, b = at((0, 0, 0)), at((0, 10, 0)), at((10, 0, 0))
    assert sorted(m.fix) == sorted([3 * o, 3 * o + 1, 3 * o + 2, 3 * a, 3 * a + 2, 3 * b + 2])
    assert len(m.loadfaces) - 1 == 24 and set(np.round(m.pressure[1:], 6)) == {0.0, 10.0}
    box = _BoxShape(10.0, 10.0, 10.0)
    assert list(box.vertex("Vertex2")) == [0, 0, 0] and list(box.vertex("Vertex4")) == [0, 10, 0]
    assert list(box.vertex("Vertex6")) == [10, 0, 0] and list(box.vertex("Vertex1")) == [0, 0, 10]


def test_fcstd_reader_resolves_planar_faces_from_stored_samples():
    """BASELINE config 2: constraints on faces of a Part::Extrusion, found through the points / normals FreeCAD
    stores with each constraint."""
    path = "/root/reference/freeCAD files/Embankment_with_Ditch_Example.FCStd"
    if not os.path.isfile(path):
        pytest.skip("reference models not present on this machine")
    from fcvm_workbench_b200.fcstd import read_fcstd, _SampledPlanarFaces
    m = read_fcstd(path)
    assert (m.ne, m.nn) == (659, 1418)
    xyz = m.nocoord
    mask = np.zeros(3 * m.nn, bool)
    mask[list(m.fix)] = True
    mask = mask.reshape(-1, 3)
    clamped = (np.abs(xyz[:, 0]) < 1e-6) | (np.abs(xyz[:, 0] - 16000) < 1e-6) | (np.abs(xyz[:, 2]) < 1e-6)
    ends = (np.abs(xyz[:, 1]) < 1e-6) | (np.abs(xyz[:, 1] + 1000) < 1e-6)
    assert mask[clamped].all()                                   # fixed faces: all three dofs
    assert mask[ends & ~clamped][:, 1].all() and not mask[ends & ~clamped][:, [0, 2]].any()   # plane strain: uy only
    assert not mask[~clamped & ~ends].any()
    # curved faces are refused, not guessed
    t = np.linspace(0, np.pi / 2, 7)
    pts = np.c_[np.cos(t), np.sin(t), 0 * t]
    with pytest.raises(NotImplementedError):
        _SampledPlanarFaces(np.vstack([pts, pts + [0, 0, 1]]), np.vstack([pts, pts]))
