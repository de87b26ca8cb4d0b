#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest8.log 2>&1; tail -5 gpurun_out/r2_pytest8.log
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],4), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]], "fallback", d["scene_stats"]["fallback_tiles"], "cand/tile", round(d["scene_stats"]["candidates_per_tile"],1), "kbar", round(d["scene_stats"]["kbar"],2))
'
for tc in 1; do
timeout 600 python bench.py --config surface_1m_1080p --steps 32 --warmup 4 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_surf8.err | python -c "$fmt" surface_new >> gpurun_out/r2_ab8.log
RTGS_PROBE_SHORT=1 timeout 600 python scratch/realistic_probe.py > gpurun_out/r2_surf8_probe.log 2>&1
(cd rt-gaussian-splat-renderer_b200/build/r1 && RTGS_PROBE_SHORT=1 timeout 600 python scratch/realistic_probe.py > ../../../gpurun_out/r2_surf8_probe_r1.log 2>&1)
done
cat gpurun_out/r2_ab8.log; tail -4 gpurun_out/r2_surf8_probe.log; tail -4 gpurun_out/r2_surf8_probe_r1.log; tail -3 gpurun_out/r2_surf8.err
