#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest6.log 2>&1; tail -15 gpurun_out/r2_pytest6.log
timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench6.log 2> gpurun_out/r2_bench6.err; python - <<'PY'
import json
for l in open('gpurun_out/r2_bench6.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['pipelined_value'],d['e2e']['sync_value'],d['kernels'], d['tiles']['ms_per_frame'], d['tiles']['ms_per_frame_latency'])
PY
tail -3 gpurun_out/r2_bench6.err
