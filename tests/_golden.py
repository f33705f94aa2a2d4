"""Loaders for tests/golden/*.npz (written by oracle/gen_golden.py from the unmodified reference)."""
import os

import numpy as np

from fcvm_workbench_b200.control import Control
from fcvm_workbench_b200.model import Model

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ANALYSES = ("tensile", "vm_uniaxial_tension", "simple_shear", "embankment", "cube2_platen", "cube2_force", "cube2_gnly", "cube2_elastic", "cube2_maxrestarts")
ORACLE_ONLY = ()                        # (branches the oracle restates but the CUDA path does not cover: none left)
BUCKLING = ("column_buckling",)         # eigen-analysis + imperfection: compared like the oracle is (see the tests)


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def model_of(z) -> Model:
    ne = len(z["m_elNodes"])
    fix = {int(d): float(v) for d, v in zip(z["m_fix_dof"], z["m_fix_val"])}
    return Model(name=str(z["m_name"]), elNodes=z["m_elNodes"], nocoord=z["m_nocoord"], fix=fix,
                 fixdof=z["m_fixdof"], movdof=z["m_movdof"],
                 materialbyElement=np.tile(z["m_materialbyElement"][0], (ne, 1)), noce=z["m_noce"],
                 loadfaces=z["m_loadfaces"], pressure=z["m_pressure"], loadvertices=z["m_loadvertices"],
                 vertexloads=z["m_vertexloads"], loadedges=z["m_loadedges"], edgeloads=z["m_edgeloads"],
                 loadfaces_uni=z["m_loadfaces_uni"], faceloads=z["m_faceloads"])


def control_of(z) -> Control:
    kw = {}
    for f in Control.__dataclass_fields__:
        v = z["c_" + f]
        kw[f] = v.item() if v.dtype.kind in "fiu" else str(v)
    return Control(**kw)


def clicks_of(z):
    out = []
    for s in z["clicks"]:
        s = str(s)
        if not s:
            continue
        if ":" in s:
            a, b = s.split(":")
            out.append((a, float(b)))
        else:
            out.append(s)
    return out


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rel_plot(k, got, z, sel=slice(None)):
    """Relative error of a per-step history against fixture ``z``.  The pressure and the triaxiality at the
    max-csr Gauss point are round-off noise in shear-dominated states (p ~ 1e-13 next to svm ~ 1e2), so they
    are measured against the stress scale / against 1 instead of against their own (vanishing) magnitude."""
    a = np.asarray(got, dtype=np.float64)[sel]
    b = np.asarray(z["r_" + k], dtype=np.float64)[sel]
    scale = np.abs(b).max(initial=0.0)
    if k == "pplot":
        scale = max(scale, np.abs(np.asarray(z["r_svmplot"], dtype=np.float64)[sel]).max(initial=0.0))
    elif k == "triaxplot":
        scale = max(scale, 1.0)
    return float(np.abs(a - b).max(initial=0.0) / max(scale, 1e-300))


def logged_iters(messages):
    """Newton iterations per load step as the reference prints them ("Step: n" / "Iteration: k, Error: e",
    fcVM.py:1315, 1344, 1455) -- the same reading ``oracle/ref_harness.iterations_per_step`` applied when the
    fixtures were written: a restart that converges at once leaves the count of the failed attempt, and a step
    abandoned after MAXIMUM RESTARTS still shows up."""
    its, cur = [], None
    for msg in messages:
        if msg.startswith("Step:") and "Load level" not in msg:
            if cur is not None:
                its.append(cur)
            cur = 0
        elif msg.startswith("Iteration:"):
            cur = int(msg.split(",")[0].split(":")[1])
    if cur is not None:
        its.append(cur)
    return its
