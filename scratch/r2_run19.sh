#!/bin/bash
L=$PWD/rt-gaussian-splat-renderer_b200/lib
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b19a.err | python -c "$fmt" tex1d_8 >> gpurun_out/r2_ab19.log
RTGS_B200_LIB=$L/lib_geotex.so timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b19b.err | python -c "$fmt" tex1d_8_geotex >> gpurun_out/r2_ab19.log
cat gpurun_out/r2_ab19.log; tail -2 gpurun_out/r2_b19b.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest19.log 2>&1; tail -3 gpurun_out/r2_pytest19.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tiles"
ncu --set full --clock-control none --import-source on -k regex:'k_shade_tiles' --launch-skip 8 --launch-count 1 -f -o gpurun_out/r02b_shade $B > gpurun_out/r02b_ncu.log 2>&1; ls -la gpurun_out/r02b_shade.ncu-rep
