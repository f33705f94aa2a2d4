"""One pass over every hot kernel at the bench size, for `ncu --set full -k regex:^k_` (see profiles/README.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from fcvm_workbench_b200 import fcVM

n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
m, c = bench.workload(n)
eng = fcVM.Engine(m.elNodes, m.nocoord, m.materialbyElement, m.fix)
if os.environ.get('PROFILE_DEFLATION'):
    eng.set_deflation(int(os.environ['PROFILE_DEFLATION']))
glv = eng.vec()
eng.assemble(glv)                                    # k_elem_stiffness, k_coo_reduce, k_apply_constraints, ...
rng = np.random.default_rng(0)
x, y = eng.vec(host=rng.normal(size=eng.ndof)), eng.vec()
eng.spmv(x, y)                                       # k_spmv_sell
du, q = eng.vec(host=2e-3 * rng.normal(size=eng.ndof)), eng.vec()
eng.gp_fill(fcVM.SIG_YIELD, 240.0)
eng.update_stress_load(None, du, q, 0.0)             # k_stress_update, k_node_gather
eng.update_peeq_csr(0.25, 0.0)                       # k_peeq_csr
eng.solve(x, y, rtol=1e-30, max_iter=int(os.environ.get('PROFILE_ITERS', '2')), raise_on_noconv=False)   # k_pcg_*
eng.synchronize()
print("ok", eng.launch_count())
