#!/bin/bash
# the default bench line (with e2e and the CPU baseline) on one GPU, then optional extra args
mkdir -p gpurun_out
timeout 900 python bench.py "$@" > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
    print("ms/step", round(d["ms_per_step"],2), "value", d["value"], "e2e", d["e2e"] and (round(d["e2e"]["ms_per_step"],1), d["e2e"]["value"]), "pcg its", d["pcg_iterations_per_step"])
    print("roofline", d["roofline"])
    for k,v in d["kernels"].items(): print("  ",k,v)
    print("standalone", d["kernels_standalone"])
    print("check", {k:(v if k not in ("newton_residual_trace",) else v[:4]) for k,v in d["check"].items()})
    print("cpu", d["cpu_baseline"] and (d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"]), "clocks", d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_full.err").read()[-3000:])
P
