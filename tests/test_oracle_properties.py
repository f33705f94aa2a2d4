"""Invariants of the oracle that do not come from golden vectors: what any correct restatement of the
reference's element routines (fcVM.py:620-816 calcGSM, 2196-2492 update_stress_load + radial return) must
satisfy on ANY mesh.  The same properties are what the GPU parity tests fall back on at sizes the oracle cannot
reach, so they are pinned here on the CPU first, on a distorted mesh with curved (non-affine) elements.
"""
import numpy as np
import pytest
import scipy.sparse as scsp

from fcvm_workbench_b200.mesh import cube_model
from fcvm_workbench_b200.model import empty_loads

E_MOD, NU = 210000.0, 0.3


def _mesh(seed=1, n=3):
    m = cube_model(n, size=4.0, mode="platen", top_disp=0.05, nxyz=(n, n + 1, n))
    rng = np.random.default_rng(seed)
    m.nocoord = m.nocoord + rng.uniform(-0.04, 0.04, m.nocoord.shape)          # mid-side nodes off the chords too
    m.materialbyElement = np.tile([E_MOD, NU, 7.8e-9], (m.ne, 1))
    return m, rng


def _stress_update(oracle, m, du, sy, sig_old=None, Et_E=0.0, disp_new=None, LD=False):
    ne, nn = m.ne, m.nn
    sig_old = np.zeros(24 * ne) if sig_old is None else sig_old
    new, test, q, pgp = np.zeros(24 * ne), np.zeros(24 * ne), np.zeros(3 * nn), np.full(4 * ne, False)
    oracle.update_stress_load(None, m.elNodes, m.nocoord, m.materialbyElement, np.full(4 * ne, sy),
                              np.zeros(3 * nn) if disp_new is None else disp_new, du, sig_old, new, test, q, Et_E, LD, pgp)
    return new, test, q, pgp


def _mises(sig):
    s = sig.reshape(-1, 6)
    p = s[:, :3].mean(axis=1)
    d = s[:, :3] - p[:, None]
    return np.sqrt(1.5 * ((d ** 2).sum(axis=1) + 2.0 * (s[:, 3:] ** 2).sum(axis=1)))


def _unconstrained_stiffness(oracle, m):
    lo = empty_loads()
    stm, row, col = oracle.calcGSM(m.elNodes, m.nocoord, m.materialbyElement, {}, 0.0, 0.0, 0.0, lo["loadfaces"],
                                   lo["pressure"], lo["loadvertices"], lo["vertexloads"], lo["loadedges"], lo["edgeloads"],
                                   lo["loadfaces_uni"], lo["faceloads"])[:3]
    low = scsp.csc_matrix((stm, (row, col)), shape=(3 * m.nn, 3 * m.nn))
    return (low + scsp.tril(low, k=-1).T).tocsr()


def _rigid_modes(xyz):
    z, o = np.zeros(len(xyz)), np.ones(len(xyz))
    x, y, w = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    modes = [np.c_[o, z, z], np.c_[z, o, z], np.c_[z, z, o], np.c_[z, -w, y], np.c_[w, z, -x], np.c_[-y, x, z]]
    return [v.ravel() for v in modes]


def test_stiffness_is_symmetric_positive_semidefinite_with_six_rigid_body_modes(oracle):
    m, rng = _mesh()
    K = _unconstrained_stiffness(oracle, m)
    scale = abs(K).max()
    for v in _rigid_modes(m.nocoord):                                          # translations and infinitesimal rotations
        assert np.abs(K @ v).max() < 1e-9 * scale * np.abs(v).max()
    for _ in range(4):
        v = rng.normal(size=3 * m.nn)
        assert v @ (K @ v) > 0.0
    d = K.toarray()
    ev = np.linalg.eigvalsh(d)
    assert (np.abs(ev) < 1e-9 * ev.max()).sum() == 6 and ev.min() > -1e-9 * ev.max()


def test_elastic_internal_force_is_the_stiffness_times_the_displacement(oracle):
    """calcGSM and update_stress_load integrate the same B^T D B: below yield q(du) = K du, linear in du."""
    m, rng = _mesh(2)
    K = _unconstrained_stiffness(oracle, m)
    du = rng.normal(0, 1e-4, 3 * m.nn)
    new, test, q, pgp = _stress_update(oracle, m, du, sy=1e9)
    assert not pgp.any() and np.abs(new - test).max() < 1e-12 * np.abs(test).max()   # recomposed p + deviator
    ref = K @ du
    assert np.abs(q - ref).max() < 1e-10 * np.abs(ref).max()
    q2 = _stress_update(oracle, m, 2.5 * du, sy=1e9)[2]
    assert np.abs(q2 - 2.5 * q).max() < 1e-12 * np.abs(q).max()


def test_rigid_translation_gives_no_stress_and_internal_forces_are_self_equilibrated(oracle):
    m, rng = _mesh(3)
    t = np.tile([0.3, -0.2, 0.5], m.nn)
    new, _, q, pgp = _stress_update(oracle, m, t, sy=100.0)
    assert np.abs(new).max() < 1e-6 and np.abs(q).max() < 1e-6 and not pgp.any()
    du = rng.normal(0, 3e-3, 3 * m.nn)                                          # well into the plastic range
    new, _, q, pgp = _stress_update(oracle, m, du, sy=100.0)
    assert pgp.mean() > 0.5
    f = q.reshape(-1, 3)
    assert np.abs(f.sum(axis=0)).max() < 1e-9 * np.abs(f).max()                 # no net force
    mom = np.cross(m.nocoord, f).sum(axis=0)
    assert np.abs(mom).max() < 1e-9 * np.abs(f).max() * np.abs(m.nocoord).max()   # no net moment (symmetric stress)


@pytest.mark.parametrize("Et_E", [0.0, 0.05])
def test_radial_return_lands_on_the_yield_surface_and_keeps_the_pressure(oracle, Et_E):
    """fcVM.py:2468-2492: the trial stress is scaled back along its deviator; the hydrostatic part is untouched,
    elastic points keep the trial stress, with perfect plasticity every plastic point ends exactly on the yield
    surface and with hardening outside it by Et/(E...) of the overshoot."""
    m, rng = _mesh(4)
    sy = 120.0
    du = rng.normal(0, 2.5e-4, 3 * m.nn)                                        # a mix of elastic and plastic points
    new, test, _, pgp = _stress_update(oracle, m, du, sy=sy, Et_E=Et_E)
    assert 0.2 < pgp.mean() < 1.0
    vm_new, vm_test = _mises(new), _mises(test)
    el = ~pgp
    assert np.abs(new.reshape(-1, 6)[el] - test.reshape(-1, 6)[el]).max() < 1e-12 * np.abs(test).max()
    assert (vm_test[el] <= sy * (1 + 1e-8)).all()
    assert (vm_test[pgp] > sy * (1 - 1e-8)).all()
    p_new, p_test = new.reshape(-1, 6)[:, :3].mean(axis=1), test.reshape(-1, 6)[:, :3].mean(axis=1)
    assert np.abs(p_new - p_test).max() < 1e-9 * np.abs(p_test).max()
    # the deviators stay parallel
    dn = new.reshape(-1, 6).copy()
    dn[:, :3] -= p_new[:, None]
    dt = test.reshape(-1, 6).copy()
    dt[:, :3] -= p_test[:, None]
    ratio = vm_new / vm_test
    assert np.abs(dn[pgp] - ratio[pgp, None] * dt[pgp]).max() < 1e-9 * vm_test.max()
    if Et_E == 0.0:
        assert np.abs(vm_new[pgp] - sy).max() < 1e-9 * sy
    else:
        assert (vm_new[pgp] > sy).all() and (vm_new[pgp] < vm_test[pgp]).all()
        # linear hardening: overshoot reduced by the same factor at every plastic point
        fac = (vm_new[pgp] - sy) / (vm_test[pgp] - sy)
        assert np.ptp(fac) < 1e-9 and 0.0 < fac[0] < 1.0


def test_stress_update_is_idempotent_for_a_zero_increment(oracle):
    """A converged state pushed through the update again with du = 0 comes back unchanged (stresses on or inside
    the yield surface stay where they are)."""
    m, rng = _mesh(5)
    du = rng.normal(0, 2e-3, 3 * m.nn)
    new, _, q, _ = _stress_update(oracle, m, du, sy=120.0)
    again, test, q2, pgp2 = _stress_update(oracle, m, np.zeros(3 * m.nn), sy=120.0, sig_old=new)
    assert np.abs(again - new).max() < 1e-9 * np.abs(new).max()
    assert np.abs(q2 - q).max() < 1e-9 * np.abs(q).max()
