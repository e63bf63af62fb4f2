#!/bin/bash
# r02 v16 on one 8 x B200 box: N = 1 reference lines, then 8 GPUs with PRN-major shards and with weighted row ranges
T=r02rows8
bash profiles/run_scaling.sh $T 1 "1" --no-parity
bash profiles/run_scaling.sh $T 1 "8" --plan prn --no-parity
bash profiles/run_scaling.sh $T 1 "8" --plan rows
bash profiles/run_scaling.sh $T 2 "1" --no-parity
bash profiles/run_scaling.sh $T 2 "8" --plan prn --no-parity
bash profiles/run_scaling.sh $T 2 "8" --plan rows
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/$T/nvidia_smi_after.csv
grep -h "sharding" gpurun_out/$T/*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], d['n_gpus'], d['run']['sharding'])"
