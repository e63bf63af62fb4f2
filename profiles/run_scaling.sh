#!/bin/bash
# usage: profiles/run_scaling.sh <tag> <config>   -- N = 1,2,4,8 back to back on one box
TAG=$1; C=${2:-2}; O=gpurun_out/$TAG; mkdir -p $O
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then python bench.py --config $C --steps 50 --warmup 5 --no-cpu > $O/scale_c${C}_n$N.json 2> $O/scale_c${C}_n$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --config $C --gpus $N --steps 50 --warmup 5 > $O/scale_c${C}_n$N.json 2> $O/scale_c${C}_n$N.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("$O/scale_c${C}_n$N.json").read().strip().splitlines()[-1])
    print("N=$N config $C: %.2f G cells/s  %.3f ms/step  e2e %.2f G  latency %.3f ms"%(d["value"]/1e9,d["ms_per_step"],d["e2e"]["value"]/1e9,d["config"]["latency_ms_32prn"]))
except Exception as e:
    print("N=$N failed", e); print(open("$O/scale_c${C}_n$N.err").read()[-1500:])
PY
done
