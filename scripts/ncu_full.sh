#!/bin/bash
# ncu --set full of selected kernels in one pass over the hot kernels: ncu_full.sh <kernel-regex> <out> [n] [count]
set -u
K=$1; OUT=$2; N=${3:-55}; C=${4:-6}
export PROFILE_DEFLATION=6144 PROFILE_ITERS=${PROFILE_ITERS:-2}
python scripts/profile_kernels.py $N > gpurun_out/plain_$OUT.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$OUT.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -c $C -f -o gpurun_out/$OUT python scripts/profile_kernels.py $N > gpurun_out/ncu_$OUT.log 2>&1 < /dev/null
tail -2 gpurun_out/ncu_$OUT.log
