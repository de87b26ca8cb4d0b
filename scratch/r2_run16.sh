#!/bin/bash
L=$PWD/rt-gaussian-splat-renderer_b200/lib
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
for i in 1 2; do
timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b16a.err | python -c "$fmt" default >> gpurun_out/r2_ab16.log
RTGS_SH_TEX=1 RTGS_B200_LIB=$L/lib_tex.so timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b16b.err | python -c "$fmt" sh_tex >> gpurun_out/r2_ab16.log
done
cat gpurun_out/r2_ab16.log; tail -2 gpurun_out/r2_b16b.err
RTGS_SH_TEX=1 RTGS_B200_LIB=$L/lib_tex.so timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -x -q 2>&1 | tail -2
