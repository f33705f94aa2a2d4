"""Multi-GPU parity check (run under torchrun, one rank per GPU):
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py
Every rank runs the element-partitioned load-stepping analysis; the curves must match the oracle's
single-domain run (1e-6) with the same Newton iterations per step, shared nodes bit-identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faulthandler

import numpy as np
import torch

faulthandler.dump_traceback_later(int(os.environ.get('FCVM_HANG_S', '100')), exit=True)   # a hang prints where
import torch.distributed as dist

from fcvm_workbench_b200 import fcVM, partition
from fcvm_workbench_b200.control import Control
from fcvm_workbench_b200.mesh import cube_model
from oracle import fcvm_oracle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
# the peer-memory halo against the NCCL interface sum on a random vector (bit-identical shared nodes)
m = cube_model(4, size=8.0, mode="platen", nxyz=(4, 4, 6), top_disp=0.05)
part = partition.slab_partition(m, world)
lm = part.local_model(rank)
comm = partition.Comm(part, rank, world)
with fcVM.Engine(lm.elNodes, lm.nocoord, lm.materialbyElement, lm.fix, device=local, comm=comm) as eng:
    if getattr(comm, "p2p", False):
        from fcvm_workbench_b200._lib import call
        import ctypes
        rng = np.random.default_rng(100 + rank)
        vh = rng.normal(size=eng.ndof)
        a, b = eng.vec(host=vh), eng.vec(host=vh)
        eng.interface_sum(a)
        for _ in range(3):                                   # epochs advance, buffers alternate
            eng.put(b, vh)
            call("fcvm_p2p_interface_sum", eng._ctx, ctypes.c_void_p(b))
        ha, hb = eng.get(a), eng.get(b)
        same = bool(np.abs(ha - hb).max() <= 1e-13 * np.abs(ha).max())
        glob = part.gather_nodal(comm.allgather(hb))
        mine = glob.reshape(-1, 3)[part.nodes[rank]].ravel()
        ident = bool(np.array_equal(mine, hb))               # every holder of a shared node has the same bits
        ok &= same and ident
        if rank == 0:
            print(f"p2p halo: world={world} equals NCCL sum={same} shared nodes bit-identical={ident}", flush=True)
    elif rank == 0:
        print("p2p halo: not attached (NCCL path)", flush=True)
for mode, kw, defl in (("platen", dict(top_disp=0.05), 0), ("force", dict(top_disp=300.0), 0), ("platen", dict(top_disp=0.05), 72)):
    m = cube_model(4, size=8.0, mode=mode, nxyz=(4, 4, 6), **kw)
    c = Control(sig_yield=240.0, nstep=6, error_max=1e-5, target_LF=1.5, Et_E=0.02, grav_z=-9.81 if mode == "force" else 0.0)
    ref = fcvm_oracle.calcDisp(m, c)
    part = partition.slab_partition(m, world)
    lm = part.local_model(rank)
    comm = partition.Comm(part, rank, world)
    o = fcVM.calcDisp(lm, c, device=local, rtol=1e-11, comm=comm, deflation=defl)
    rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))
    same_iters = list(o["iters"]) == list(ref["iters"])
    errs = {k: rel(o[k], ref[k]) for k in ("lout", "un", "peeqplot", "csrplot")}
    disp = part.gather_nodal(comm.allgather(o["displacements"]))
    sig = part.gather_gauss(comm.allgather(o["stresses"]))
    errs["displacements"] = rel(disp, ref["displacements"])
    errs["stresses"] = rel(sig, ref["stresses"])
    good = same_iters and all(v < 1e-6 for k, v in errs.items() if k in ("lout", "un", "peeqplot", "csrplot")) and \
        errs["displacements"] < 1e-5 and errs["stresses"] < 1e-5
    ok &= good
    if rank == 0:
        print(f"{mode} deflation={defl} p2p={getattr(comm, 'p2p', False)}: world={world} iters {list(o['iters'])} same={same_iters} "
              + " ".join(f"{k}={v:.1e}" for k, v in errs.items()) + (" OK" if good else " FAIL"), flush=True)
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
