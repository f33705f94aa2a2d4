"""bench.py's contract on a machine without a GPU: the reference arm prints one JSON line with the agreed
keys; the product arm refuses to run (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "2",
                        "--cpu-n", "4"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "newton_gauss_point_updates_per_s" and d["unit"] == "GP-updates/s"
    assert d["steps"] == 3 and d["warmup"] == 2 and d["higher_is_better"] is True and d["value"] > 0
    # the unmodified reference when it can be loaded here (numba + /root/reference or oracle/_ref), else the port
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--n", "2"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)


def test_weak_scaling_meshes_keep_the_elements_per_rank():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.weak_cells(55, 1) == (55, 55, 55)
    assert bench.weak_cells(55, 2) == (55, 55, 110)
    assert bench.weak_cells(55, 4) == (55, 110, 110)
    assert bench.weak_cells(55, 8) == (110, 110, 110)
    for w in (1, 2, 4, 8):
        nx, ny, nz = bench.weak_cells(7, w)
        assert nx * ny * nz == w * 7 ** 3
    with pytest.raises(SystemExit):
        bench.weak_cells(55, 3)
    # the weak mesh keeps the element size and the nominal strain of the cube
    m1, _ = bench.workload(3)
    m2, _ = bench.workload(3, bench.weak_cells(3, 2))
    assert m2.ne == 2 * m1.ne
    h1 = m1.nocoord[:, 0].max() / 3
    assert abs(m2.nocoord[:, 0].max() / 3 - h1) < 1e-12 and abs(m2.nocoord[:, 2].max() / 6 - h1) < 1e-12
    top1 = max(v for v in m1.fix.values())
    top2 = max(v for v in m2.fix.values())
    assert abs(top2 / m2.nocoord[:, 2].max() - top1 / m1.nocoord[:, 2].max()) < 1e-15


def test_check_object_compares_with_the_committed_single_gpu_trace(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import types

    import numpy as np

    import bench
    ref = json.load(open(os.path.join(ROOT, "profiles", "check_n55.json")))
    a = types.SimpleNamespace(n=55, write_check=False, rtol=1e-8, steps=10, warmup=3)
    sw = types.SimpleNamespace(errors=list(ref["newton_residual_trace"][:13]))
    out = dict(iters=np.array(ref["newton_iters_per_step"]), lout=np.array(ref["lout"]), un=np.array(ref["un"]))
    chk = bench.sweep_check(a, 4, None, sw, out)
    assert chk["vs_single_gpu"]["ok"] and chk["vs_single_gpu"]["residual_trace_rel_diff"] == 0.0
    out["lout"] = out["lout"] * (1 + 1e-5)
    assert not bench.sweep_check(a, 4, None, sw, out)["vs_single_gpu"]["ok"]
