#!/bin/bash
# usage: scratch/ab.sh libA libB ...  (names under lib/), prints value/kernel per lib, twice
for i in 1 2; do for l in "$@"; do
RTGS_B200_LIB=$PWD/rt-gaussian-splat-renderer_b200/lib/$l python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for x in sys.stdin:
    if x.startswith('{'):
        d=json.loads(x); print('$l', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['kernel_ms'],4))
" >> gpurun_out/ab.log
done; done
