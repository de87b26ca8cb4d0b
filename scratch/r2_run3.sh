#!/bin/bash
L=$PWD/rt-gaussian-splat-renderer_b200/lib
run() {  # tag, env...
  tag=$1; shift
  env "$@" RTGS_DEBUG_OCC=1 timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_bench_$tag.err | python -c "
import sys,json
for x in sys.stdin:
    if x.startswith('{'):
        d=json.loads(x); print('$tag', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'sync', round(d['e2e']['sync_value'],1), 'ms', round(d['ms_per_step'],4), [(k['kernel'], round(k['ms'],4)) for k in d['kernels']])
" >> gpurun_out/r2_ab3.log
  grep "rtgs:" gpurun_out/r2_bench_$tag.err | sort | uniq >> gpurun_out/r2_ab3.log
}
for i in 1; do
run v1_m0 RTGS_B200_LIB=$L/lib_v1.so RTGS_RENDER_MODE=0
run v2_m0 RTGS_B200_LIB=$L/lib_v2.so RTGS_RENDER_MODE=0
run v3_m0 RTGS_B200_LIB=$L/lib_v3.so RTGS_RENDER_MODE=0
done
cat gpurun_out/r2_ab3.log
