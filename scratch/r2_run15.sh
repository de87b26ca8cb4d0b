#!/bin/bash
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]], "fallback", d["scene_stats"]["fallback_tiles"], "cand/tile", round(d["scene_stats"]["candidates_per_tile"],1))
'
for lim in 960 1500 2500 4000 7000; do
RTGS_HEAVY_LIMIT=$lim timeout 600 python bench.py --config surface_1m_1080p --steps 32 --warmup 4 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_surf15_$lim.err | python -c "$fmt" surface_limit_$lim >> gpurun_out/r2_ab15.log
done
cat gpurun_out/r2_ab15.log; tail -3 gpurun_out/r2_surf15_7000.err
RTGS_HEAVY_LIMIT=2500 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest15.log 2>&1; tail -3 gpurun_out/r2_pytest15.log
