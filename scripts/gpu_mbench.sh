#!/bin/bash
# bench on N GPUs (torchrun launched by bench.py itself): gpu_mbench.sh N [bench args...]
N=$1; shift
mkdir -p gpurun_out
FCVM_HANG_S=${FCVM_HANG_S:-200} timeout 500 python bench.py --gpus $N "$@" > gpurun_out/mb_$N.json 2> gpurun_out/mb_$N.err; echo "bench N=$N rc=$?"
python - gpurun_out/mb_$N <<'P'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f+".json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], d["scaling"], "elements", d["config"]["elements"], "ms/step", round(d["ms_per_step"],2), "value", round(d["value"]/1e6,2), "M  pcg its", d["pcg_iterations_per_step"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1))
    for k,v in d["kernels"].items(): print("  ",k,{a:b for a,b in v.items() if a in ("avg_ms","launches","share","frac_of_hbm_peak")})
    c=d["check"]; print("check", c.get("vs_single_gpu"), c["newton_iters_per_step"], c["lout"], c["un"])
    print("setup", d.get("setup_s"), "exchange", d["config"].get("exchange"))
except Exception as e:
    print("failed", e); print(open(f+".err").read()[-3000:])
P
