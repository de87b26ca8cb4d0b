#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "just_outside or inside_the_cloud" > gpurun_out/r2_pytest46a.log 2>&1; tail -5 gpurun_out/r2_pytest46a.log
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b46.err | python -c "$fmt" near_t_refine >> gpurun_out/r2_ab46.log
timeout 600 python bench.py --config surface_1m_1080p --steps 48 --warmup 5 --no-cpu-baseline --no-tiles 2>> gpurun_out/r2_b46.err | python -c "$fmt" near_t_refine_surface >> gpurun_out/r2_ab46.log
cat gpurun_out/r2_ab46.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest46.log 2>&1; tail -3 gpurun_out/r2_pytest46.log
