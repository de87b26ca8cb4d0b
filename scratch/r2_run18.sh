#!/bin/bash
L=$PWD/rt-gaussian-splat-renderer_b200/lib
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest18.log 2>&1; tail -3 gpurun_out/r2_pytest18.log
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "serial", round(d["serial_value"],1), "e2e", round(d["e2e"]["value"],1), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b18a.err | python -c "$fmt" tex2d_8 >> gpurun_out/r2_ab18.log
RTGS_B200_LIB=$L/lib_tex0.so timeout 600 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b18b.err | python -c "$fmt" ldg_only >> gpurun_out/r2_ab18.log
timeout 600 python bench.py --config surface_1m_1080p --steps 32 --warmup 4 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b18c.err | python -c "$fmt" surface_tex >> gpurun_out/r2_ab18.log
timeout 600 python bench.py --config 3m_deg3_2160p --steps 16 --warmup 3 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b18d.err | python -c "$fmt" cfg4_tex >> gpurun_out/r2_ab18.log
RTGS_B200_LIB=$L/lib_tex0.so timeout 600 python bench.py --config 3m_deg3_2160p --steps 16 --warmup 3 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b18e.err | python -c "$fmt" cfg4_ldg >> gpurun_out/r2_ab18.log
timeout 600 python bench.py --config 100k_deg0_1080p --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_b18f.err | python -c "$fmt" cfg2 >> gpurun_out/r2_ab18.log
cat gpurun_out/r2_ab18.log; tail -2 gpurun_out/r2_b18a.err
