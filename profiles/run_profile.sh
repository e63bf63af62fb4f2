#!/bin/bash
# Run on the GPU box: bench lines, ncu launch list, ncu full capture of the dominant kernel.
# usage: profiles/run_profile.sh <tag> [extra bench args]
set -u
TAG=${1:-r01}; shift || true
EXTRA="$@"
O=gpurun_out/$TAG
mkdir -p $O
python bench.py $EXTRA > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 3000 $O/bench_c2.json
python bench.py --config 1 --no-cpu $EXTRA > $O/bench_c1.json 2> $O/bench_c1.err; tail -c 1500 $O/bench_c1.json
for C in 2 1; do
  CMD="python bench.py --config $C --steps 2 --warmup 1 --no-cpu $EXTRA"
  $CMD > $O/plain_c$C.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_c$C.csv $CMD > $O/ncu_launch_c$C.log 2>&1
  $CMD > $O/plain2_c$C.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -o $O/search_c$C $CMD > $O/ncu_full_c$C.log 2>&1
  tail -3 $O/ncu_full_c$C.log
done
ls -la $O
