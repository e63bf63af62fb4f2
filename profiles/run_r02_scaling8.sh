#!/bin/bash
# r02 multi-GPU measurements on one 8-GPU box (gpurun --gpus 8): what BASELINE.json asks at N > 1.
T=r02s8
bash profiles/run_scaling.sh $T 1 "1 2 4 8"
bash profiles/run_scaling.sh $T 1 "8" --xchg nccl
bash profiles/run_scaling.sh $T 2 "1 8"
bash profiles/run_scaling.sh $T 3 "1 8" --no-parity
bash profiles/run_scaling.sh $T 4 "8"
bash profiles/run_scaling.sh $T 4 "8" --sweep-shard epochs
bash profiles/run_scaling.sh $T 5 "8" --no-parity
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/$T/nvidia_smi_after.csv
