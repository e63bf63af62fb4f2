#!/bin/bash
# r02 final library: GPU tests, default bench lines of configs 1 and 2 (CPU baseline + parity check inside)
O=gpurun_out/r02final; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
for C in 1 2; do
  python bench.py --config $C > $O/bench_c$C.json 2> $O/bench_c$C.err
  python -c "
import json;d=json.loads(open('$O/bench_c$C.json').read().strip().splitlines()[-1]);print($C, 'value %.3f G'%(d['value']/1e9), 'ms %.3f'%d['ms_per_step'], 'e2e %.3f G'%(d['e2e']['value']/1e9), 'lat %.3f'%d['run']['latency_ms_32prn'], 'frac %.4f'%d['roofline']['frac'], 'k2 %.3f'%d['roofline']['kernel_ms'], d['parity_checked']['ok'], d['clocks'])"
done
