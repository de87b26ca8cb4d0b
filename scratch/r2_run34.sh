#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest34.log 2>&1; tail -3 gpurun_out/r2_pytest34.log
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), "pipe", round(d["e2e"]["pipelined_value"],1), "sync", round(d["e2e"]["sync_value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b34.err | python -c "$fmt" after_heavy_lists >> gpurun_out/r2_ab34.log
cat gpurun_out/r2_ab34.log
