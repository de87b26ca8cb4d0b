#!/bin/bash
L=$PWD/rt-gaussian-splat-renderer_b200/lib
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]], "fallback", d["scene_stats"]["fallback_tiles"], d["scene_stats"]["stack_high_water"])
'
for v in f00 f10 f08 f14; do
RTGS_B200_LIB=$L/lib_$v.so timeout 600 python bench.py --config surface_1m_1080p --steps 32 --warmup 4 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_surf14_$v.err | python -c "$fmt" surface_$v >> gpurun_out/r2_ab14.log
done
timeout 600 python bench.py --config surface_1m_1080p --steps 32 --warmup 4 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_surf14_def.err | python -c "$fmt" surface_default_f18 >> gpurun_out/r2_ab14.log
cat gpurun_out/r2_ab14.log
RTGS_STATS_FALLBACK_ONLY=1 timeout 300 python scratch/surface_stats.py 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest14.log 2>&1; tail -3 gpurun_out/r2_pytest14.log
