#!/bin/bash
# usage: scripts/ncu_kernel.sh <kernel-regex> <out-name> [n] [count]   (run under gpurun; one GPU)
set -u
K=$1; OUT=$2; N=${3:-40}; C=${4:-1}
python scripts/profile_kernels.py $N > gpurun_out/plain_$OUT.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$OUT.log; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$K" -c $C -f -o gpurun_out/$OUT python scripts/profile_kernels.py $N > gpurun_out/ncu_$OUT.log 2>&1 < /dev/null
tail -1 gpurun_out/ncu_$OUT.log
