#!/bin/bash
# r02 final (v14 kernels): 8-GPU lines with and without the IF relay, configs 1 and 2; config 3 and 5 once more
T=r02s8b
bash profiles/run_scaling.sh $T 1 "1 8"
bash profiles/run_scaling.sh $T 1 "8" --no-relay
bash profiles/run_scaling.sh $T 2 "1 8"
bash profiles/run_scaling.sh $T 2 "8" --no-relay
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/$T/nvidia_smi_after.csv
