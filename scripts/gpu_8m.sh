#!/bin/bash
# BASELINE config 4 on N GPUs: strong sweep point of the 8M mesh (n = 110) and weak sweep point (6*55^3 per rank)
N=$1
mkdir -p gpurun_out
FCVM_HANG_S=500 bash scripts/gpu_mbench.sh $N --steps 10 --warmup 3 --no-e2e --no-cpu --cells 110 --scaling strong; cp gpurun_out/mb_$N.json gpurun_out/r02_bench_8M_strong_N$N.json
if [ "$N" -gt 1 ]; then
  FCVM_HANG_S=400 bash scripts/gpu_mbench.sh $N --steps 10 --warmup 3 --no-e2e --scaling weak; cp gpurun_out/mb_$N.json gpurun_out/r02_bench_weak_N$N.json
fi
