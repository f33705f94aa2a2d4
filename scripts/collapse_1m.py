"""The north-star run: full load-displacement curve of the 1M-element von Mises collapse analysis.
   python scripts/collapse_1m.py [n]                                 (one GPU)
   torchrun --nproc-per-node N ... scripts/collapse_1m.py [n]        (N GPUs, element-partitioned)
Writes gpurun_out/collapse_n{n}_N{world}.json (curve, Newton iterations per step, wall time); the curves of
different N must agree (compare with scripts/compare_curves.py)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from fcvm_workbench_b200 import fcVM, partition

n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
rtol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
deflation = int(sys.argv[3]) if len(sys.argv) > 3 else bench.DEFLATION
rank, world, local = bench.dist_env()
torch.cuda.set_device(local)
comm = None
m, ctl = bench.workload(n)
if world > 1:
    import faulthandler
    import torch.distributed as dist
    faulthandler.dump_traceback_later(int(os.environ.get("FCVM_HANG_S", "600")), exit=True)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    part = partition.slab_partition(m, world)
    comm = partition.Comm(part, rank, world)
    lm = part.local_model(rank)
else:
    lm = m
t0 = time.time()
eng = fcVM.Engine(lm.elNodes, lm.nocoord, lm.materialbyElement, lm.fix, device=local, comm=comm)
eng.synchronize()
t1 = time.time()
out = fcVM.calcDisp(lm, ctl, engine=eng, rtol=rtol, max_iter=200000, deflation=deflation)
eng.synchronize()
t2 = time.time()
if rank == 0:
    res = dict(n=n, elements=m.ne, nodes=m.nn, world=world, rtol=rtol, deflation_grid=eng.deflation_grid, setup_s=t1 - t0, analysis_s=t2 - t1,
               newton_iterations=int(out["iterat_tot"]), iters=[int(i) for i in out["iters"]],
               pcg_iterations=int(np.sum(out["pcg_iterations"])), lout=[float(v) for v in out["lout"]],
               un=[float(v) for v in out["un"]], lbd=[float(v) for v in out["lbd"]],
               peeqplot=[float(v) for v in out["peeqplot"]], csrplot=[float(v) for v in out["csrplot"]],
               nplastic=[int(v) for v in out["nplastic"]], launches=int(out["launches"]))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/collapse_n{n}_N{world}_d{deflation}.json", "w") as f:
        json.dump(res, f)
    print(f"n={n} N={world}: {m.ne} elements, {len(res['iters'])} load steps, {res['newton_iterations']} Newton iterations, "
          f"{res['pcg_iterations']} PCG iterations, analysis {res['analysis_s']:.1f} s (setup {res['setup_s']:.1f} s); "
          f"final reaction {res['lout'][-1]:.6e}, plastic Gauss points {res['nplastic'][-1]} of {4 * m.ne}", flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
