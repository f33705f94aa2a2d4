#!/bin/bash
# multi-GPU checks on N GPUs: parity script (p2p and NCCL paths), then short benches.  usage: gpu_mgpu.sh N
N=${1:-2}
mkdir -p gpurun_out
for p in ${P2PS:-1 0}; do
  FCVM_P2P=$p FCVM_HANG_S=200 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py > gpurun_out/m_check_p2p$p.log 2>&1; echo "check p2p=$p rc=$?"
  grep -E "OK|FAIL|p2p halo|rror" gpurun_out/m_check_p2p$p.log | tail -8
done
for p in ${P2PS:-1 0}; do
  FCVM_P2P=$p FCVM_HANG_S=250 timeout 500 python bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/m_bench_p2p$p.json 2> gpurun_out/m_bench_p2p$p.err; echo "bench p2p=$p rc=$?"
  python - gpurun_out/m_bench_p2p$p <<'P'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f+".json").read().strip().splitlines()[-1])
    print("ms/step", round(d["ms_per_step"],2), "pcg its", d["pcg_iterations_per_step"], {k:(v.get("avg_ms"),v.get("launches")) for k,v in d["kernels"].items()})
except Exception as e:
    print("failed", e, open(f+".err").read()[-1500:])
P
done
