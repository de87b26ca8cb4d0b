#!/bin/bash
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), "sync", round(d["e2e"]["sync_value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b56.err | python -c "$fmt" final >> gpurun_out/r2_ab56.log
timeout 600 python bench.py --config surface_1m_1080p --steps 48 --warmup 5 --no-cpu-baseline --no-tiles 2>> gpurun_out/r2_b56.err | python -c "$fmt" final_surface >> gpurun_out/r2_ab56.log
cat gpurun_out/r2_ab56.log
