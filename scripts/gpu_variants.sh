#!/bin/bash
# experiments: block shapes of the fused PCG kernel (variant libraries built by hand, see profiles/README.md)
mkdir -p gpurun_out
for v in "$@"; do
  FCVM_LIB_PATH=$PWD/fcvm_workbench_b200/libfcvm_var_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err; echo "$v rc=$?"
done
python - "$@" <<'P'
import json,sys
for n in sys.argv[1:]:
    try:
        d=json.loads(open(f"gpurun_out/v_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "pcg its", d["pcg_iterations_per_step"], "phases", d.get("pcg_phases_ms_per_iteration"))
    except Exception as e:
        print(n, "failed", e, open(f"gpurun_out/v_{n}.err").read()[-500:])
P
