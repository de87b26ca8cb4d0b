#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_pytest13.log 2>&1; tail -3 gpurun_out/r2_pytest13.log; grep -i "high water\|tree depth\|ksat mode" gpurun_out/r2_pytest13.log
timeout 600 python bench.py > gpurun_out/r2_bench13.log 2> gpurun_out/r2_bench13.err; python - <<'PY'
import json
for l in open('gpurun_out/r2_bench13.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'serial',d['serial_value'],'e2e',d['e2e']['value'],d['e2e']['pipelined_value'],d['e2e']['sync_value'],'compact',d['e2e']['compact']['value']); print(d['kernels']); print(d['roofline']['frac'], d['scene_stats']['stack_high_water'], d['cpu_baseline']); print({k:v for k,v in d['tiles'].items() if k not in ('by_mode','single_gpu_by_mode','limiting_kernel')})
PY
tail -3 gpurun_out/r2_bench13.err
