#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest10.log 2>&1; tail -8 gpurun_out/r2_pytest10.log
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],4), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]], "fallback", d["scene_stats"]["fallback_tiles"])
'
RTGS_HEAVY_OVERLAP=0 timeout 600 python bench.py --config surface_1m_1080p --steps 32 --warmup 4 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_surf10a.err | python -c "$fmt" surface_seq >> gpurun_out/r2_ab10.log
timeout 600 python bench.py --config surface_1m_1080p --steps 32 --warmup 4 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_surf10b.err | python -c "$fmt" surface_overlap >> gpurun_out/r2_ab10.log
timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b10.err | python -c "$fmt" bench >> gpurun_out/r2_ab10.log
cat gpurun_out/r2_ab10.log; tail -3 gpurun_out/r2_surf10b.err
timeout 300 python scratch/surface_stats.py 2>&1 | tail -3
