#!/bin/bash
# final single-GPU evidence: the default bench line (with e2e and the reference CPU baseline), the reference arm,
# the ncu launch list of the bench command, ncu --set full of the top kernels
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_N1.json 2> gpurun_out/r02_bench_N1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "reference arm rc=$?"
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_launch.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_n55.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1 < /dev/null; echo "ncu launches rc=$?"
PROFILE_ITERS=3 bash scripts/ncu_full.sh "k_elastic_apply_affine|k_gather_apply|k_pcg_step_bulk|k_coarse_rhs|k_gemv_bulk|k_expand|k_stress_update_pair|k_node_gather|k_spmv_sell|k_elem_stiffness|k_coo_reduce" r02_full_final 55 16
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_bench_N1.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],2), "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "cpu", d["cpu_baseline"])
print(open("gpurun_out/r02_bench_reference_arm.json").read()[:600])
P
