#!/bin/bash
# r02 v16: row-range shards with a calibrated root share against PRN-major shards, 2 GPUs, configs 1 and 2
T=r02rows2
bash profiles/run_scaling.sh $T 1 "2" --plan prn --no-parity
bash profiles/run_scaling.sh $T 1 "2" --plan rows
bash profiles/run_scaling.sh $T 2 "2" --plan prn --no-parity
bash profiles/run_scaling.sh $T 2 "2" --plan rows
grep -h "sharding" gpurun_out/$T/*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], d['run']['sharding'])"
