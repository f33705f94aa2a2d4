"""Headless fcVM analysis on the GPU: control file + mesh in, ``.out`` / ``.vtk`` out.

    python -m fcvm_workbench_b200 --inp "control files/tensile.inp" --fcstd "freeCAD files/tensile.FCStd"
    python -m fcvm_workbench_b200 --inp run.inp --npz model.npz --clicks add,add --out results/
    python -m fcvm_workbench_b200 --inp run.inp --cube 20 --mode platen
    python -m fcvm_workbench_b200 --inp run.inp --cube 55 --gpus 8        # one process per GPU (torchrun), element slabs

What the workbench macro does between "run" and the result files (fcVM.FCMacro:100-262), minus
FreeCAD: setUpInput -> calcGSM -> calcDisp -> mapStresses -> .out + .vtk.  ``--clicks`` scripts the
buttons of the load-displacement window (add / rev / stop); without it the analysis stops after
the first ``nstep`` load steps, like pressing "stop".
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m fcvm_workbench_b200", description=__doc__.split("\n\n")[0])
    ap.add_argument("--inp", required=True, help="fcVM control file (21 lines)")
    src = ap.add_mutually_exclusive_group(required=True)
    src.add_argument("--fcstd", help="FreeCAD document holding the mesh, material and constraints")
    src.add_argument("--npz", help="model bundle written by model.Model.save_npz")
    src.add_argument("--cube", type=int, help="synthetic structured cube with this many cells per edge")
    ap.add_argument("--mode", default="platen", help="boundary conditions of --cube (tension, platen, punch, force)")
    ap.add_argument("--top-disp", type=float, default=0.05)
    ap.add_argument("--clicks", default="", help="comma-separated: add, rev, stop, add:<target load factor>")
    ap.add_argument("--out", default=".", help="directory for <name>.out and <name>.vtk")
    ap.add_argument("--rtol", type=float, default=1e-10, help="PCG relative residual of every linear solve")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1, help="GPUs of this box: the mesh is partitioned element-wise, one "
                                                        "process per GPU (launched through torch.distributed.run)")
    ap.add_argument("--partition", choices=("auto", "slab", "compact"), default="auto",
                    help="with --gpus: contiguous ranges of the element list as it is (slab) or after coordinate "
                         "bisection (compact); auto = slab for --cube, compact for meshes read from a file")
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args(argv)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if a.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # started plainly: launch the ranks the way the driver launches bench.py
        import socket
        import subprocess
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", str(port), "-m", "fcvm_workbench_b200"] + list(argv if argv is not None else sys.argv[1:])
        return subprocess.call(cmd)
    if a.gpus != world:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE is {world}")

    from . import fcVM, results
    from .control import read_control
    from .model import Model
    ctl = read_control(a.inp)
    t0 = time.time()
    if a.fcstd:
        from .fcstd import read_fcstd
        m = read_fcstd(a.fcstd)
        name = os.path.splitext(os.path.basename(a.fcstd))[0]
    elif a.npz:
        m = Model.load_npz(a.npz)
        name = m.name
    else:
        from .mesh import cube_model
        m = cube_model(a.cube, mode=a.mode, top_disp=a.top_disp)
        name = m.name
    clicks = []
    for c in filter(None, a.clicks.split(",")):
        clicks.append((c.split(":")[0], float(c.split(":")[1])) if ":" in c else c)
    say = (lambda *s: None) if (a.quiet or rank != 0) else (lambda *s: print(*s, flush=True))
    say(f"{name}: {m.ne} elements, {m.nn} nodes ({time.time() - t0:.2f} s input)" + (f", {world} GPUs" if world > 1 else ""))
    t0 = time.time()
    deflation = fcVM.AUTO_DEFLATION if m.nn >= 20000 else 0
    averaged = ctl.averaged_option == "averaged"
    if world > 1:
        # one process per GPU: element slabs, shared nodes completed over NVLink; rank 0 gathers and writes
        import torch
        import torch.distributed as dist
        from . import partition
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # structured cubes list their elements slab by slab; anything read from a file is renumbered by
        # coordinate bisection first, so that a rank has a few neighbours instead of all of them
        how = a.partition if a.partition != "auto" else ("compact" if (a.fcstd or a.npz) else "slab")
        part = (partition.compact_partition if how == "compact" else partition.slab_partition)(m, world)
        comm = partition.Comm(part, rank, world)
        res = fcVM.calcDisp(part.local_model(rank), ctl, clicks=clicks, device=local, rtol=a.rtol, log=say, comm=comm,
                            deflation=deflation)
        t1 = time.time()
        for key in ("displacements", "disp_el"):
            res[key] = part.gather_nodal(comm.allgather(res[key]))
        for key in ("stresses", "peeq", "sigmises", "csr"):
            res[key] = part.gather_gauss(comm.allgather(res[key]))
        # shared nodes carry the completed load on every rank that holds them: sum the gathered global vector
        res["crip"] = part.original_gauss_point(res["crip"])
        res["loadsum"] = tuple(part.gather_nodal(comm.allgather(res["glv"])).reshape(-1, 3).sum(axis=0))
        dist.destroy_process_group()
        if rank != 0:
            return 0
        a.device = local
    else:
        with fcVM.Engine(m.elNodes, m.nocoord, m.materialbyElement, m.fix, device=a.device) as eng:
            res = fcVM.calcDisp(m, ctl, clicks=clicks, engine=eng, rtol=a.rtol, log=say, deflation=deflation)
        t1 = time.time()
    with fcVM.Engine(m.elNodes, m.nocoord, m.materialbyElement, m.fix, device=a.device) as eng:      # mapStresses
        for which, key in ((fcVM.SIG_NEW, "stresses"), (fcVM.PEEQ, "peeq"), (fcVM.SIGMISES, "sigmises"), (fcVM.CSR, "csr")):
            eng.gp_put(which, res[key])
        t10 = eng.map_stresses(averaged, ctl.sig_yield, m.noce)
    os.makedirs(a.out, exist_ok=True)
    x = fcVM.gauss_point_coordinates(m.elNodes, m.nocoord)
    results.write_out(os.path.join(a.out, name + ".out"), name, m.ne, m.nn, ctl.gnl, ctl.nstep, res["loadsum"], res, x=x,
                      eigenval=res.get("eigenval"))
    results.write_vtk(os.path.join(a.out, name + ".vtk"), m.elNodes, m.nocoord, res["displacements"], *t10)
    say(f"load-stepping {t1 - t0:.2f} s, {res['iterat_tot']} Newton iterations, {int(np.sum(res['pcg_iterations']))} PCG "
        f"iterations; wrote {name}.out and {name}.vtk to {a.out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
