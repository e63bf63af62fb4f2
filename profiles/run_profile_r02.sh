#!/bin/bash
# r02: ncu launch lists (every launch with its device time) and one full capture of the search kernel, configs 1 and 2.
O=gpurun_out/r02p14; mkdir -p $O
for C in 1 2; do
  CMD="python bench.py --config $C --steps 2 --warmup 1 --no-cpu --no-parity"
  $CMD > $O/plain_c$C.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_c$C.csv $CMD > $O/ncu_launch_c$C.log 2>&1
  $CMD > $O/plain2_c$C.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -o $O/search_c$C $CMD > $O/ncu_full_c$C.log 2>&1
  tail -2 $O/ncu_full_c$C.log
done
ls -la $O
