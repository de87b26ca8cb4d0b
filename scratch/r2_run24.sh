#!/bin/bash
L=$PWD/rt-gaussian-splat-renderer_b200/lib
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest24.log 2>&1; tail -3 gpurun_out/r2_pytest24.log
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b24_$tag.err | python -c "$fmt" $tag >> gpurun_out/r2_ab24.log; }
run default X=1
run lists_3ctas RTGS_B200_LIB=$L/lib_k1c3.so
cat gpurun_out/r2_ab24.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
