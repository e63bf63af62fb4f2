#!/bin/bash
# usage: profiles/ncu_one.sh <tag> <config> [extra bench args]  -- one ncu --set full capture of search_kernel
TAG=$1; C=$2; shift 2; EXTRA="$@"
O=gpurun_out/$TAG; mkdir -p $O
CMD="python bench.py --config $C --steps 2 --warmup 1 --no-cpu $EXTRA"
$CMD > $O/plain_c$C.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -o $O/search_c$C $CMD > $O/ncu_full_c$C.log 2>&1
tail -2 $O/ncu_full_c$C.log; tail -c 600 $O/plain_c$C.log
