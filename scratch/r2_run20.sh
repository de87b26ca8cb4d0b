#!/bin/bash
L=$PWD/rt-gaussian-splat-renderer_b200/lib
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b20_$tag.err | python -c "$fmt" $tag >> gpurun_out/r2_ab20.log; }
run tex8_runtime_fallback X=1
run tex8_compile_time RTGS_B200_LIB=$L/lib_texct.so
run tex8_geotex RTGS_B200_LIB=$L/lib_geotex.so
run ldg_only RTGS_B200_LIB=$L/lib_tex0.so
cat gpurun_out/r2_ab20.log; tail -2 gpurun_out/r2_b20_tex8_geotex.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest20.log 2>&1; tail -3 gpurun_out/r2_pytest20.log
