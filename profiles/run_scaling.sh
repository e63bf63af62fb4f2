#!/bin/bash
# usage: profiles/run_scaling.sh <tag> <config> "<N list>" [extra bench args]  -- bench.py at several GPU counts on one box
TAG=$1; C=${2:-1}; NS=${3:-"1 2 4 8"}; shift 3; EXTRA="$@"; O=gpurun_out/$TAG; mkdir -p $O
for N in $NS; do
  F=$O/scale_c${C}_n$N$(echo "$EXTRA" | tr -d ' -')
  if [ $N -eq 1 ]; then timeout 900 python bench.py --config $C --steps 20 --warmup 5 --no-cpu $EXTRA > $F.json 2> $F.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --config $C --gpus $N --steps 20 --warmup 5 $EXTRA > $F.json 2> $F.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("$F.json").read().strip().splitlines()[-1])
    print("N=$N config $C $EXTRA: %.2f G cells/s  %.3f ms/step  e2e %.2f G  latency %.3f ms  parity %s  phases %s"%(d["value"]/1e9,d["ms_per_step"],d["e2e"]["value"]/1e9,d["run"]["latency_ms_32prn"],d.get("parity_checked",{}).get("ok"),d["run"].get("exchange_phases_ms")))
except Exception as e:
    print("N=$N failed", e); print(open("$F.err").read()[-2500:])
PY
done
