#!/bin/bash
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b57.err | python -c "$fmt" rare_second_chunk >> gpurun_out/r2_ab57.log
cat gpurun_out/r2_ab57.log
timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "dense_cluster or heavy or render_paths" 2>&1 | tail -2
