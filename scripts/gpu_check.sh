#!/bin/bash
# parity suite, then quick bench variants:  gpu_check.sh "ENV=.." "ENV=.." ...
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -15 gpurun_out/c_pytest.log
bash scripts/gpu_quick.sh "$@"
