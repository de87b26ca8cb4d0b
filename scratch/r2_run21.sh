#!/bin/bash
L=$PWD/rt-gaussian-splat-renderer_b200/lib
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b21_$tag.err | python -c "$fmt" $tag >> gpurun_out/r2_ab21.log; }
run w10_default X=1
run w8_128regs RTGS_B200_LIB=$L/lib_w8.so
run w11_80regs RTGS_B200_LIB=$L/lib_w11.so
run tex10 RTGS_B200_LIB=$L/lib_t10.so
cat gpurun_out/r2_ab21.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tiles"
ncu --set full --clock-control none --import-source on -k regex:'k_shade_tiles' --launch-skip 8 --launch-count 1 -f -o gpurun_out/r02b_shade $B > gpurun_out/r02b_ncu.log 2>&1; ls -la gpurun_out/r02b_shade.ncu-rep
