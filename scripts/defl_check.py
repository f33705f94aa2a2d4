"""PCG with and without rigid-body-mode deflation: iterations, time, agreement (one GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from fcvm_workbench_b200 import fcVM

n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
targets = [int(t) for t in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 768, 3072]
m, c = bench.workload(n)
sols = []
for tgt in targets:
    eng = fcVM.Engine(m.elNodes, m.nocoord, m.materialbyElement, m.fix)
    grid = eng.set_deflation(tgt)
    glv = eng.vec()
    eng.timer_start(); eng.assemble(glv); ms_a = eng.timer_stop_ms()
    f, zero, x = eng.vec(), eng.vec(), eng.vec()
    eng.residual(1.0, glv, zero, f); eng.axpby(1.0, eng.buf(fcVM.MODF), 1.0, f)
    for tol in (1e-8,):
        eng.timer_start(); its, rr = eng.solve(f, x, tol, 100000); ms = eng.timer_stop_ms()
        eng.timer_start(); its, rr = eng.solve(f, x, tol, 100000); ms = eng.timer_stop_ms()
    # true residual
    y = eng.vec(); eng.spmv(x, y); eng.axpby(1.0, f, -1.0, y)
    tr = eng.norm(y) / eng.norm(f)
    sols.append(eng.get(x))
    print(f"n={n} ne={m.ne} deflation target {tgt} grid {grid}: assemble {ms_a:.1f} ms, pcg {its} its, {ms:.1f} ms "
          f"({ms / max(its, 1):.3f} ms/it), relres {rr:.2e}, true relres {tr:.2e}, "
          f"diff to first {np.abs(sols[-1] - sols[0]).max() / np.abs(sols[0]).max():.1e}", flush=True)
    eng.close()
