#!/bin/bash
# launch list (device time of every launch) of one pass over the hot kernels: ncu_launches.sh <out-name> [n]
set -u
OUT=$1; N=${2:-55}
export PROFILE_DEFLATION=6144 PROFILE_ITERS=6
python scripts/profile_kernels.py $N > gpurun_out/plain_$OUT.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$OUT.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/$OUT.csv python scripts/profile_kernels.py $N > gpurun_out/ncu_$OUT.log 2>&1 < /dev/null
python - gpurun_out/$OUT.csv <<'P'
import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    if r[ui]=="ns": v/=1e3
    elif r[ui]=="ms": v*=1e3
    k=r[ki].split("(")[0]
    agg.setdefault(k,[]).append(v)
for k,v in agg.items():
    print(f"{k[:70]:70s} n={len(v):4d} last={v[-1]:9.1f} us  min={min(v):9.1f}  mean={sum(v)/len(v):9.1f}")
P
