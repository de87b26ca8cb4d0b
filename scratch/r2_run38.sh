#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_deep_tree.py tests/test_golden_reference.py -m gpu -x -q > gpurun_out/r2_pytest38.log 2>&1; tail -4 gpurun_out/r2_pytest38.log
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
timeout 600 python bench.py --config surface_1m_1080p --steps 48 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b38.err | python -c "$fmt" two_level_fused >> gpurun_out/r2_ab38.log
cat gpurun_out/r2_ab38.log
