#!/bin/bash
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
for hs in 0 2; do
RTGS_HEAVY_SLAB=$hs RTGS_SLAB_RANK=24 timeout 600 python bench.py --config surface_1m_1080p --steps 48 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b33.err | python -c "$fmt" heavy_slab_$hs >> gpurun_out/r2_ab33.log
done
cat gpurun_out/r2_ab33.log
