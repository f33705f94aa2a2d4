"""Exploratory timing of the main kernels on one GPU (not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fcvm_workbench_b200 import fcVM
from fcvm_workbench_b200.mesh import cube_model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rtol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-9
t0 = time.time(); m = cube_model(n, size=10.0, mode="platen", top_disp=0.1); t1 = time.time()
print(f"n={n} ne={m.ne} nn={m.nn} mesh {t1-t0:.1f}s", flush=True)
eng = fcVM.Engine(m.elNodes, m.nocoord, m.materialbyElement, m.fix); eng.synchronize(); t2 = time.time()
print(f"engine setup {t2-t1:.2f}s  stats {eng.matrix_stats()}", flush=True)
glv = eng.vec()
eng.timer_start(); eng.assemble(glv); ms = eng.timer_stop_ms(); print(f"assemble {ms:.2f} ms")
eng.timer_start(); eng.assemble(glv); ms = eng.timer_stop_ms(); print(f"assemble(2nd) {ms:.2f} ms")
eng.profile(True); eng.assemble(glv); p = eng.profile_get(); eng.profile(False)
print("assemble parts:", {k: round(v[0], 3) for k, v in p.items() if v[1]})
x, y = eng.vec(host=np.random.default_rng(0).normal(size=eng.ndof)), eng.vec()
for _ in range(3): eng.spmv(x, y)
eng.timer_start()
for _ in range(20): eng.spmv(x, y)
ms = eng.timer_stop_ms() / 20
st = eng.matrix_stats()
alg = st["blocks_real"] * 76 + eng.ndof * 8 * 2
print(f"spmv {ms:.3f} ms  stored-bytes {st['bytes']/1e9:.2f} GB -> {st['bytes']/ms/1e6:.0f} GB/s stored, algorithmic {alg/ms/1e6:.0f} GB/s")
du, q = eng.vec(host=1e-3*np.random.default_rng(1).normal(size=eng.ndof)), eng.vec()
eng.gp_fill(fcVM.SIG_YIELD, 240.0)
for _ in range(3): eng.update_stress_load(None, du, q, 0.0)
eng.profile(True)
for _ in range(10): eng.update_stress_load(None, du, q, 0.0)
p = eng.profile_get(); eng.profile(False)
su, ng = p["stress_update"][0]/10, p["node_gather"][0]/10
print(f"stress_update {su:.3f} ms ({4*m.ne/su/1e6:.1f} G GP/s, {m.ne*(40+192+32+192+192+4+240+100)/su/1e6:.0f} GB/s alg)  node_gather {ng:.3f} ms")
# elastic solve
f = eng.vec(); zero = eng.vec()
eng.residual(1.0, glv, zero, f); eng.axpby(1.0, eng.buf(fcVM.MODF), 1.0, f)
ue = eng.vec()
for tol in (1e-6, 1e-8, rtol):
    eng.timer_start(); its, rr = eng.solve(f, ue, tol, 100000); ms = eng.timer_stop_ms()
    print(f"pcg rtol={tol:g}: {its} its, relres {rr:.2e}, {ms:.1f} ms, {ms/its:.3f} ms/it", flush=True)
eng.profile(True); its, rr = eng.solve(f, ue, 1e-6, 100000); p = eng.profile_get(); eng.profile(False)
print({k: (round(v[0],2), v[1]) for k, v in p.items()}, "its", its)
