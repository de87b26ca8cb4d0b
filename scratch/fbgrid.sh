#!/bin/bash
for g in 1 37 296; do
RTGS_FB_GRID=$g python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for x in sys.stdin:
    if x.startswith('{'):
        d=json.loads(x); print('fbgrid $g', round(d['value'],1), round(d['e2e']['value'],1), [round(k['ms'],4) for k in d['kernels']])
" >> gpurun_out/fbgrid.log
done
