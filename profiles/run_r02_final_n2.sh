#!/bin/bash
# r02 v16 sanity at 2 GPUs: the default (weighted row ranges + rebalance) on the sweep, the 401-bin and the 2001-bin grids
T=r02fin2
for C in 4 3 5; do
  F=gpurun_out/$T/c$C; mkdir -p gpurun_out/$T
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600+C)) bench.py --config $C --gpus 2 --steps 4 --warmup 3 > $F.json 2> $F.err
  python - <<PY
import json
try:
    d=json.loads(open("$F.json").read().strip().splitlines()[-1])
    print("config $C N=2: %.2f G cells/s %.3f ms/step parity %s | %s"%(d["value"]/1e9,d["ms_per_step"],d.get("parity_checked",{}).get("ok"),d["run"]["sharding"]))
except Exception as e:
    print("config $C failed", e); print(open("$F.err").read()[-2000:])
PY
done
