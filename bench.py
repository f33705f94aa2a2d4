#!/usr/bin/env python
"""Throughput of the fcVM Newton / load-stepping hot path on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

Workload (``config.workload``): synthetic structured C3D10 cube, ``6 n^3`` elements (n = 55:
998,250 elements, 1,367,631 nodes), clamped bottom, rough rigid platen pushed into the top
face, elastic-perfectly-plastic von Mises material, displacement-controlled load sweep
(fcVM.py:1304-1559 semantics).  A *step* is one Newton iteration of that sweep: one linear
solve with the elastic stiffness (PCG on the device instead of CHOLMOD's ``factor(f)``), the
arc-length update, the radial-return stress update at every Gauss point with the internal-force
assembly, and the residual norm -- plus the per-load-step bookkeeping that falls between two
iterations (update_PEEQ_CSR, state roll-over).  Warm-up iterations are the first W Newton
iterations of the sweep (the elastic first load step needs none and is passed before).

``value`` = Gauss points x Newton iterations / second with every array resident in HBM
(``newton_iters_per_s`` and the raw stress-update rate are given beside it); ``e2e`` is the same
sweep driven through the HOST-buffer C ABI (hostpath.HostEngine: numpy arrays in page-locked
memory, every heavy call pays its PCIe copies).  What one PCG iteration streams (1.9 GB: element geometry,
element vectors, work vectors, coarse operators) and the Gauss-point state of a Newton iteration (0.96 GB) are
far larger than the 126 MB L2, so no explicit L2 flush is needed between steps.

For N > 1 (torchrun) the same 1M-element mesh is partitioned element-wise into N slabs (``--scaling strong``,
the default: BASELINE's "1M elements at 1/2/4/8 B200"); ``--scaling weak`` keeps 6 n^3 elements PER RANK instead (the
mesh grows to n x n x 2n, n x 2n x 2n, 2n x 2n x 2n cells at 2/4/8 ranks: 8M elements at N = 8, BASELINE config 4) and
``--n 110 --scaling strong`` is the strong sweep of the 8M mesh.  Inside the PCG iteration the ranks exchange through
mapped peer memory (NVLink); everything else goes through NCCL.  Every line carries a ``check`` object (Newton
residual trace, load and displacement history of the sweep); with a committed single-GPU trace of the same
sweep (``profiles/check_*.json``) the run fails when its history differs by more than 1e-6.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFLATION = 6144          # coarse unknowns of the deflation level (10 x 10 x 10 boxes x 6 modes on the cube)
METRIC = "newton_gauss_point_updates_per_s"
UNIT = "GP-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--n", "--cells", dest="n", type=int, default=55, help="cells per cube edge (elements = 6 n^3)")
    ap.add_argument("--rtol", type=float, default=1e-8, help="PCG relative residual per linear solve")
    ap.add_argument("--cpu-n", type=int, default=14, help="cube edge of the bounded CPU sample")
    ap.add_argument("--deflation", type=int, default=DEFLATION,
                    help="unknowns of the rigid-body-mode coarse level of the PCG preconditioner (0 = block-Jacobi only)")
    ap.add_argument("--scaling", default="strong", choices=("strong", "weak"),
                    help="N > 1: partition the 6 n^3 mesh (strong) or keep 6 n^3 elements per rank (weak)")
    ap.add_argument("--write-check", action="store_true", help="N = 1: write profiles/check_<sweep>.json")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def weak_cells(n, world):
    """cells per direction of the weak-scaling mesh: 6 n^3 elements per rank, doubled direction by direction"""
    f = [1, 1, 1]
    k, d = world, 2
    while k > 1:
        if k % 2:
            raise SystemExit("--scaling weak needs a power-of-two number of ranks")
        f[d] *= 2
        d = (d - 1) % 3
        k //= 2
    return n * f[0], n * f[1], n * f[2]


def workload(n, nxyz=None):
    from fcvm_workbench_b200.control import Control
    from fcvm_workbench_b200.mesh import cube_model
    if nxyz is None:
        m = cube_model(n, size=10.0, mode="platen", top_disp=0.05)
    else:
        # same element size and the same nominal strain as the cube: the platen moves in proportion to the height
        m = cube_model(n, size=10.0 * nxyz[0] / n, mode="platen", top_disp=0.05 * nxyz[2] / n, nxyz=nxyz)
    c = Control(sig_yield=240.0, nstep=10, iterat_max=20, error_max=1e-3, relax=1.2, target_LF=1.0, Et_E=0.0)
    return m, c


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 - 0.1 <= t <= t1 + 0.3):
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_traffic(kernel, n, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture of the same
    kernel at the same size (profiles/traffic.json); None when there is no capture of this configuration."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            t = json.load(f).get(kernel)
    except (OSError, ValueError):
        return None
    if t and t.get("n") == n and t.get("n_gpus") == world:
        return int(t["dram_bytes_per_launch"])
    return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class Sweep:
    """Drives calcDisp and times Newton iterations W+1 .. W+K with CUDA events on the engine's stream."""

    def __init__(self, eng, W, K, barrier=None, profile_stride=0):
        self.eng, self.W, self.K, self.barrier, self.stride = eng, W, K, barrier, profile_stride
        self.ms = None
        self.t0 = self.t1 = None
        self.launch0 = self.launch1 = 0
        self.pcg_its = []
        self.errors = []
        self.bytes0 = (0, 0)
        self.prof = None

    def _sync(self):
        self.eng.synchronize()
        if self.barrier is not None:
            self.barrier()
            self.eng.synchronize()

    def hook(self, d):
        from fcvm_workbench_b200.fcVM import StopAnalysis
        it = d["iterat_tot"]
        self.errors.append(float(d["error"]))
        dev = getattr(self.eng, "dev", self.eng)
        if it == self.W:
            self._sync()
            if self.stride:
                dev.profile(self.stride)
            self.bytes0 = (getattr(self.eng, "h2d_bytes", 0), getattr(self.eng, "d2h_bytes", 0))
            self.launch0 = dev.launch_count()
            self.t0 = time.time()
            dev.timer_start()
        elif it > self.W:
            self.pcg_its.append(d["pcg_iterations"])
        if it == self.W + self.K:
            self._sync()
            self.ms = dev.timer_stop_ms()
            self.t1 = time.time()
            self.launch1 = dev.launch_count()
            self.bytes1 = (getattr(self.eng, "h2d_bytes", 0), getattr(self.eng, "d2h_bytes", 0))
            if self.stride:
                self.prof = dev.profile_get()
                dev.profile(0)
            raise StopAnalysis()


def run_sweep(model, ctl, eng, W, K, rtol, barrier=None, profile_stride=0, deflation=None):
    from fcvm_workbench_b200 import fcVM
    sw = Sweep(eng, W, K, barrier, profile_stride)
    # W = 0: the hook of iteration 0 does not exist; start the clock on the first call instead
    if W == 0:
        raise SystemExit("--warmup must be >= 1 (the contract asks for >= 3)")
    out = fcVM.calcDisp(model, ctl, engine=eng, rtol=rtol, max_iter=200000, on_iteration=sw.hook, deflation=deflation)
    if sw.ms is None:
        raise SystemExit(f"the load sweep ended after {out['iterat_tot']} Newton iterations, fewer than "
                         f"warmup+steps = {W + K}: lower --steps or raise the load")
    return sw, out


def kernel_report(eng, model, prof, hbm_peak, world=1, total_ms=None):
    """Algorithmic bytes per launch of each kernel family ('Kernels and their rooflines')."""
    st = eng.matrix_stats()
    ne, nn = eng.ne, eng.nn
    defl = eng.deflation_grid is not None
    ncl = int(np.prod(eng.deflation_grid)) if defl else 0
    alg = {
        # matrix-free product: conn 40 B + mask 4 B + stored geometry 80 B + element vector written 240 B per element;
        # gather: element vectors read 240 B + node->element list 40 B per element, x, r read and y written per node
        "product": ne * (40 + 4 + 80 + 240) + ne * (240 + 40) + nn * (3 * 24 + 3),
        # conn 40 B, sig_old 192 B + sig_yield 32 B read, sig_new + sig_test 384 B + pgp 4 B written,
        # nodal xyz + du read once (48 B per node), element force vector written (240 B)
        "stress_update": ne * (40 + 192 + 32 + 384 + 4 + 240) + nn * 48,
        # element force vectors read (240 B), node->element list (40 B), qin written (24 B per node)
        "node_gather": ne * (240 + 40) + nn * 24,
        # w, u, p, s, x, r read, p, s, x, r, u written, inverse diagonal blocks read
        "pcg_step": nn * (6 * 24 + 5 * 24 + 72),
        # K Z streamed in single precision (18 x 4 B + 4 B node index per stored entry), r, y, xyz, fixdof (24 B each)
        # and the box list (4 B) per node
        "coarse_rhs": int(eng.deflation_stats()["entries"] * 76 + nn * 100) if defl else 0,
        "coarse_product": 4 * (6 * ncl) ** 2 // max(world, 1),
        "coarse_expand": nn * (24 + 24 + 24 + 4 + 24),
    }
    flops = {"product": ne * 1900.0}
    names = {"spmv": "product"}                     # family 0 = the PCG's product (matrix-free here)
    rep = {}
    for k0, (ms, timed, seen) in prof.items():
        if timed == 0:
            continue
        k = names.get(k0, k0)
        if k == "pcg_vector":
            k = "pcg_vector_scopes"
        avg = ms / timed
        r = {"avg_ms": round(avg, 5), "timed_launches": timed, "launches": seen}
        if total_ms:
            r["share"] = round(avg * seen / total_ms, 4)
        if k in alg and alg[k]:
            gbs = alg[k] / avg / 1e6
            r.update(algorithmic_bytes=int(alg[k]), achieved_gbs=round(gbs, 1), frac_of_hbm_peak=round(gbs / hbm_peak, 4))
        if k in flops:
            r["fp64_tflops"] = round(flops[k] / avg / 1e9, 2)
        rep[k] = r
    # the vector step is what remains of the "pcg_vector" scopes after the coarse level and the exchanges
    return rep, alg


def sweep_check(a, world, nxyz, sw, out):
    """What a reader needs to compare runs across N: the Newton residual after every iteration of the sweep and the
    load / displacement history of the completed load steps -- compared with the committed single-GPU trace of the
    same sweep when there is one (rc != 0 beyond 1e-6 on the histories)."""
    chk = {"newton_residual_trace": [float(f"{e:.9e}") for e in sw.errors],
           "newton_iters_per_step": [int(i) for i in out["iters"]],
           "lout": [float(v) for v in out["lout"]], "un": [float(v) for v in out["un"]]}
    tag = f"n{a.n}" + ("" if nxyz is None else "_weak%d" % world)
    path = os.path.join(ROOT, "profiles", f"check_{tag}.json")
    if a.write_check and world == 1:
        with open(path, "w") as f:
            json.dump(dict(chk, sweep=tag, pcg_rtol=a.rtol, steps=a.steps, warmup=a.warmup), f)
    if os.path.isfile(path) and not a.write_check:
        with open(path) as f:
            ref = json.load(f)

        def dev(x, y):
            k = min(len(x), len(y))
            x, y = np.asarray(x[:k], dtype=float), np.asarray(y[:k], dtype=float)
            return float(np.abs(x - y).max(initial=0.0) / max(np.abs(y).max(initial=0.0), 1e-300)), k

        d_l, k_l = dev(chk["lout"], ref["lout"])
        d_u, k_u = dev(chk["un"], ref["un"])
        d_e, k_e = dev(chk["newton_residual_trace"], ref["newton_residual_trace"])
        k_i = min(len(chk["newton_iters_per_step"]), len(ref["newton_iters_per_step"]))
        chk["vs_single_gpu"] = {"reference": os.path.relpath(path, ROOT), "lout_rel_diff": d_l, "un_rel_diff": d_u,
                                "residual_trace_rel_diff": d_e, "points_compared": [k_l, k_u, k_e],
                                "same_newton_iters": chk["newton_iters_per_step"][:k_i] == ref["newton_iters_per_step"][:k_i],
                                "ok": bool(d_l < 1e-6 and d_u < 1e-6
                                           and chk["newton_iters_per_step"][:k_i] == ref["newton_iters_per_step"][:k_i])}
    return chk


def extra_kernel_legs(eng, hbm_peak, reps=10):
    """Kernels that the geometrically linear sweep runs once (assembly) or not at all (assembled SpMV, used by the
    large-displacement branch): ``reps`` timed launches each, CUDA events around every launch, after the sweep."""
    st = eng.matrix_stats()
    ne, nn = eng.ne, eng.nn
    glv = eng.vec()
    x, y = eng.vec(host=np.random.default_rng(0).normal(size=eng.ndof)), eng.vec()
    eng.profile(1)
    for _ in range(reps):
        eng.assemble(glv)
    for _ in range(reps):
        eng.spmv(x, y)
    prof = eng.profile_get()
    eng.profile(0)
    alg = {"elem_stiffness": ne * (40 + 3960), "coo_reduce": st["blocks_real"] * 72 + 100 * ne * 76,
           "spmv": st["blocks_real"] * 76 + 2 * 24 * nn}
    flops = {"elem_stiffness": ne * 6500.0}            # ~6.5 kFLOP per element (4 Jacobians, 4 gradient tiles, 55 blocks)
    rep = {}
    for k in ("elem_stiffness", "coo_reduce", "spmv"):
        ms, timed, seen = prof[k]
        if not timed:
            continue
        avg = ms / timed
        gbs = alg[k] / avg / 1e6
        rep[k] = {"avg_ms": round(avg, 5), "timed_launches": timed, "algorithmic_bytes": int(alg[k]),
                  "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm_peak, 4)}
        if k in flops:
            rep[k]["fp64_tflops"] = round(flops[k] / avg / 1e9, 2)
    rep["note"] = ("the sweep re-uses one assembly; spmv is the assembled block-SELL product (large-displacement branch, "
                   "start vectors), the PCG of this sweep applies the operator matrix-free")
    return rep


def roofline_of(kern, hbm_peak, peak_src, n, world):
    """The kernel family with the largest share of the timed region."""
    fam = {k: v for k, v in kern.items() if k in ("product", "pcg_step", "coarse_rhs", "coarse_product", "stress_update")}
    if not fam:
        return None
    top = max(fam, key=lambda k: fam[k]["avg_ms"] * fam[k]["launches"])
    r = fam[top]
    names = {"product": "k_elastic_apply_affine + k_gather_apply (matrix-free product w = K u of the PCG: element pass + "
                        "deterministic gather with the fused dot products)",
             "pcg_step": "k_pcg_step", "coarse_rhs": "k_coarse_rhs", "coarse_product": "k_gemv", "stress_update": "k_stress_update_pair"}
    return {"kernel": names[top], "bound": "hbm", "achieved": r.get("achieved_gbs"), "peak": hbm_peak, "unit": "GB/s",
            "frac": r.get("frac_of_hbm_peak"), "traffic": measured_traffic(top, n, world), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": r.get("algorithmic_bytes"), "avg_launch_ms": r.get("avg_ms"),
            "share_of_timed_region": r.get("share"),
            "note": ("FP64-issue bound, not HBM bound: ~1.9 kFLOP per element against 0.36 kB of algorithmic traffic; the "
                     "assembled SpMV it replaces runs at the copy bandwidth (kernels_standalone.spmv) but moves 8x the bytes")
            if top == "product" else None,
            "fp64_tflops": r.get("fp64_tflops")}


class _StdoutToStderr:
    """The reference prints its progress on stdout (also from numba-compiled code, i.e. through the C stream):
    keep this process's stdout for the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def cpu_baseline(W, K, n):
    """The reference's CPU path on a bounded sample of the same workload, timed on this box's host cores.

    kind "reference": the UNMODIFIED reference (``oracle/_ref``, put there by ``oracle/make_ref.py`` at build time,
    or /root/reference) -- its numba-jitted ``calcGSM`` / ``update_stress_load`` / ``update_PEEQ_CSR`` and its own
    ``calcDisp`` driven through ``oracle/ref_harness.py``; scikit-sparse (CHOLMOD) is not installable offline, so
    its ``cholesky`` / ``factor(b)`` are served by SuperLU (scipy) on the same matrix.  kind "port": the oracle
    (C restatement of the numba routines + the same SuperLU solve), used when numba or the copy is missing."""
    m, c = workload(n)
    threads = os.environ.get("NUMBA_NUM_THREADS")
    try:
        from oracle import ref_harness as rh
        have_ref = rh.available()
    except Exception:
        have_ref = False
    if have_ref:
        import numba
        with _StdoutToStderr():
            dt, setup = rh.time_reference(m, c, W, K)
        kind = "reference"
        what = (f"the unmodified reference (numba {numba.__version__}, its jitted element routines are serial: "
                f"NUMBA_NUM_THREADS={threads or numba.config.NUMBA_NUM_THREADS} has no effect) + its calcDisp; "
                f"CHOLMOD stand-in: one SuperLU factorisation")
    else:
        from oracle import fcvm_oracle as orc
        stamps = []

        def hook(d):
            stamps.append(time.perf_counter())
            if d["iterat_tot"] == W + K:
                raise orc.StopAnalysis()

        t_setup = time.perf_counter()
        try:
            orc.calcDisp(m, c, on_iteration=hook)
        except orc.StopAnalysis:
            pass
        if len(stamps) < W + K:
            raise SystemExit("cpu sample ended before warmup+steps Newton iterations")
        dt, setup = stamps[W + K - 1] - stamps[W - 1], stamps[0] - t_setup
        kind = "port"
        what = "the oracle port (C restatement of the numba element routines) + one SuperLU factorisation"
    return {"value": 4 * m.ne * K / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "numba_num_threads": threads,
            "host_cores": os.cpu_count(),
            "sample": f"cube n={n}: {m.ne} elements, {K} Newton iterations after {W} warm-up of the same platen sweep; "
                      f"{what} (set-up incl. factorisation {setup:.1f} s, untimed); a direct factorisation of the "
                      f"full n=55 system (4.1M dofs) does not fit the time box",
            "newton_iters_per_s": K / dt, "ms_per_step": 1e3 * dt / K, "elements": m.ne}


def main():
    a = parse()
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
    if os.environ.get("FCVM_HANG_S"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["FCVM_HANG_S"]), exit=True)
    rank, world, local = dist_env()
    if a.impl == "reference":
        if rank != 0:
            return 0
        cb = cpu_baseline(a.warmup, a.steps, a.cpu_n)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"structured C3D10 cube, von Mises platen load sweep; bounded CPU sample "
                                       f"n={a.cpu_n} ({cb['elements']} elements) of the n={a.n} workload"},
                "newton_iters_per_s": cb["newton_iters_per_s"], "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    if a.gpus > 1 and world == 1 and "WORLD_SIZE" not in os.environ:
        # started as plain `python bench.py --gpus N`: launch the ranks the way the driver does
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + \
            ["--cells" if v == "--n" else v for v in sys.argv[1:]]      # torchrun's parser takes "--n" for an abbreviation of its own
        return subprocess.call(cmd)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    if a.gpus != world:
        raise SystemExit(f"bench.py: --gpus {a.gpus} but WORLD_SIZE is {world}")
    from fcvm_workbench_b200 import fcVM
    from fcvm_workbench_b200.hostpath import HostEngine
    torch.cuda.set_device(local)
    barrier = None
    comm = None
    if world > 1:
        import torch.distributed as dist
        from fcvm_workbench_b200 import partition
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        tok = torch.zeros(1, device="cuda")

        def barrier():
            dist.all_reduce(tok)
            torch.cuda.synchronize()

    hbm_peak, peak_src = peaks()
    t_mesh = time.time()
    nxyz = weak_cells(a.n, world) if (a.scaling == "weak" and world > 1) else None
    gmodel, ctl = workload(a.n, nxyz)
    ne_total = gmodel.ne
    defl_target = a.deflation
    if a.deflation and ne_total > 1.5 * 6 * 55 ** 3:
        # boxes of the same size as on the 1M cube as far as the dense coarse inverse allows (16384 unknowns)
        defl_target = int(min(16380, a.deflation * ne_total / (6 * 55 ** 3)))
    if world > 1:
        part = partition.slab_partition(gmodel, world)
        model = part.local_model(rank)
        comm = partition.Comm(part, rank, world)
    else:
        model = gmodel
    t_mesh = time.time() - t_mesh

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    eng = fcVM.Engine(model.elNodes, model.nocoord, model.materialbyElement, model.fix, device=local, comm=comm)
    sw, out = run_sweep(model, ctl, eng, a.warmup, a.steps, a.rtol, barrier, profile_stride=31, deflation=defl_target)
    p2p = bool(getattr(comm, "p2p", False)) if comm is not None else False
    check = sweep_check(a, world, nxyz, sw, out)
    extra = extra_kernel_legs(eng, hbm_peak) if world == 1 else {}
    defl_grid = eng.deflation_grid
    ms = sw.ms
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop(sw.t0, sw.t1) if rank == 0 else None
    kern, alg = kernel_report(eng, model, sw.prof, hbm_peak, world, total_ms=sw.ms)
    launches = sw.launch1 - sw.launch0
    # raw Gauss-point update rate of the stress-update kernel + its deterministic force assembly
    gp_rate = None
    if "stress_update" in kern:
        t_su = kern["stress_update"]["avg_ms"] + kern.get("node_gather", {"avg_ms": 0.0})["avg_ms"]
        gp_rate = 4 * ne_total / (t_su * 1e-3) if world == 1 else 4 * model.ne * world / (t_su * 1e-3)
    eng.close()

    e2e = None
    if not a.no_e2e:
        hcomm = partition.Comm(part, rank, world) if world > 1 else None
        heng = HostEngine(model.elNodes, model.nocoord, model.materialbyElement, model.fix, device=local, comm=hcomm)
        hs, _ = run_sweep(model, ctl, heng, a.warmup, a.steps, a.rtol, barrier, deflation=defl_target)
        hms, hb = hs.ms, [hs.bytes1[0] - hs.bytes0[0], hs.bytes1[1] - hs.bytes0[1]]
        host_split = {k: round(v, 3) for k, v in sorted(heng.host_seconds.items(), key=lambda kv: -kv[1])}
        if world > 1:
            t = torch.tensor([hms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            hms = float(t.item())
            tb = torch.tensor(hb, device="cuda", dtype=torch.float64)
            dist.all_reduce(tb)
            hb = [float(v) for v in tb.tolist()]
        e2e = {"value": 4 * ne_total * a.steps / (hms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(hb[0] / a.steps), "d2h_bytes_per_step": int(hb[1] / a.steps),
               "ms_per_step": hms / a.steps, "newton_iters_per_s": a.steps / (hms * 1e-3),
               "host_seconds_whole_sweep": host_split,
               "path": "hostpath.HostEngine: fcvm_host_solve + fcvm_host_update_stress_load on page-locked numpy arrays"
                       + (" (per rank, bytes summed over ranks)" if world > 1 else "")}
        heng.close()

    cb = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cb = cpu_baseline(min(a.warmup, 3), min(a.steps, 10), a.cpu_n)

    if rank == 0:
        line = {
            "metric": METRIC, "value": 4 * ne_total * a.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"structured C3D10 {'cube' if nxyz is None else 'block %dx%dx%d cells' % tuple(nxyz)} "
                                   f"n={a.n}: {ne_total} elements, {gmodel.nn} nodes, von Mises "
                                   "(elastic-perfectly-plastic) platen compression, displacement-controlled load sweep",
                       "elements": ne_total, "nodes": gmodel.nn, "step": "one Newton iteration (PCG solve + arc-length "
                       "update + radial-return stress update + internal force + residual)",
                       "pcg_rtol": a.rtol, "partition": f"{world} element slab(s)" + (
                           f", {a.scaling} scaling" + (f" ({6 * a.n ** 3} elements per rank)" if a.scaling == "weak" else "")
                           if world > 1 else ""),
                       "exchange": ("peer memory over NVLink inside the PCG iteration (neighbour-only halo, rank-ordered "
                                    "sums), NCCL elsewhere" if p2p else ("NCCL" if world > 1 else "none")),
                       "product": "elastic operator recomputed element by element (matrix-free) inside the PCG",
                       "preconditioner": ("block-Jacobi + rigid-body-mode deflation, boxes %s" % (defl_grid,)
                                          if defl_grid else "block-Jacobi"),
                       "l2": "inputs exceed L2: per rank and PCG iteration %.2f GB of element geometry, element vectors, "
                             "work vectors and coarse operators, %.2f GB of Gauss-point state per Newton iteration, "
                             "vs 126 MB" % (ne_total / world * 1.9e-6, ne_total / world * 0.96e-6)},
            "newton_iters_per_s": a.steps / (ms * 1e-3),
            "stress_update_gauss_points_per_s": gp_rate,
            "pcg_iterations_per_step": float(np.mean(sw.pcg_its)) if sw.pcg_its else None,
            "gpu_launches": int(launches),
            "e2e": e2e,
            "roofline": roofline_of(kern, hbm_peak, peak_src, a.n, world),
            "kernels": kern,
            "kernels_standalone": extra,
            "check": check,
            "cpu_baseline": cb,
            "clocks": clk,
            "setup_s": {"mesh": round(t_mesh, 2)},
        }
        print(json.dumps(line))
    bad = bool(check.get("vs_single_gpu")) and not check["vs_single_gpu"]["ok"]
    if bad and rank == 0:
        print("bench.py: the history of this sweep differs from the committed single-GPU trace by more than 1e-6 "
              f"({check['vs_single_gpu']})", file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
