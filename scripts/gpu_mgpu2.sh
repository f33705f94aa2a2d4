#!/bin/bash
# parity script on N ranks (peer-memory path), then the bench line: gpu_mgpu2.sh N [bench args]
N=$1; shift
mkdir -p gpurun_out
FCVM_HANG_S=150 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py > gpurun_out/m_check_$N.log 2>&1; rc=$?; echo "check N=$N rc=$rc"
grep -E "OK|FAIL|p2p halo|rror" gpurun_out/m_check_$N.log | tail -8
if [ $rc -ne 0 ]; then tail -30 gpurun_out/m_check_$N.log; exit 1; fi
FCVM_HANG_S=200 bash scripts/gpu_mbench.sh $N "$@"
