#!/bin/bash
# GPU check of the fused PCG kernel: parity suite, then the bench with the fused and the launch-per-phase solver.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/a_bench_fused.json 2> gpurun_out/a_bench_fused.err; echo "fused rc=$?"
FCVM_PCG_CLASSIC=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/a_bench_classic.json 2> gpurun_out/a_bench_classic.err; echo "classic rc=$?"
FCVM_COARSE_FP64=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/a_bench_fused64.json 2> gpurun_out/a_bench_fused64.err; echo "fused64 rc=$?"
python - <<'P'
import json
for n in ("fused","classic","fused64"):
    try:
        d=json.loads(open(f"gpurun_out/a_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "pcg its", d["pcg_iterations_per_step"], "phases", d.get("pcg_phases_ms_per_iteration"), {k:(v.get("avg_ms"),v.get("launches")) for k,v in d["kernels"].items()})
    except Exception as e:
        print(n, "failed", e)
P
