#!/bin/bash
# two-GPU evidence: multi-GPU tests (parity, runner), then the bench line
mkdir -p gpurun_out
FCVM_HANG_S=150 timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_runner.py -x -q > gpurun_out/n2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/n2_pytest.log
FCVM_HANG_S=200 bash scripts/gpu_mbench.sh 2 --steps 20 --warmup 5
cp gpurun_out/mb_2.json gpurun_out/r02_bench_N2.json
