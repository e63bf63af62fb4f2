#!/bin/bash
# r02 final (v16): GPU tests, bench lines of configs 1 and 2 (with the CPU baseline and the parity check), ncu launch
# lists and one full capture of the search kernel per config
O=gpurun_out/r02p16; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
for C in 1 2; do
  python bench.py --config $C > $O/bench_c$C.json 2> $O/bench_c$C.err; tail -c 600 $O/bench_c$C.json
  CMD="python bench.py --config $C --steps 2 --warmup 1 --no-cpu --no-parity"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_c$C.csv $CMD > $O/ncu_launch_c$C.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -o $O/search_c$C $CMD > $O/ncu_full_c$C.log 2>&1
  tail -2 $O/ncu_full_c$C.log
done
ls -la $O
