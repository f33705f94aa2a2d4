#!/bin/bash
# quick single-GPU bench runs with environment variants:  gpu_quick.sh "VAR=1 VAR2=2" "..." ...
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 300 python bench.py --steps ${STEPS:-5} --warmup 3 --no-e2e --no-cpu > gpurun_out/q_$i.json 2> gpurun_out/q_$i.err; echo "[$v] rc=$?"
  python - "$v" gpurun_out/q_$i <<'P'
import json,sys
n,f=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f+".json").read().strip().splitlines()[-1])
    print(n, "ms/step", round(d["ms_per_step"],2), "pcg its", d["pcg_iterations_per_step"], {k:(v.get("avg_ms"),v.get("launches")) for k,v in d["kernels"].items()})
except Exception as e:
    print(n, "failed", e, open(f+".err").read()[-800:])
P
done
