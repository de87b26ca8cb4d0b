#!/bin/bash
# round 2, GPU run 1: GPU tests of the restructured kernels, then A/B of render modes / variants on the bench scene
L=$PWD/rt-gaussian-splat-renderer_b200/lib
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/r2_run1_env.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_bench_$tag.err | python -c "
import sys,json
for x in sys.stdin:
    if x.startswith('{'):
        d=json.loads(x); print('$tag', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'sync', round(d['e2e']['sync_value'],1), 'ms', round(d['ms_per_step'],4), [(k['kernel'], round(k['ms'],4)) for k in d['kernels']])
" >> gpurun_out/r2_ab1.log
}
for i in 1 2; do
run mode2_tail RTGS_RENDER_MODE=2
run mode2_notail RTGS_RENDER_MODE=2 RTGS_TAIL_LAUNCH=0
run mode0 RTGS_RENDER_MODE=0
run take16_mode2 RTGS_B200_LIB=$L/librtgs_take16.so
run take16_mode0 RTGS_B200_LIB=$L/librtgs_take16.so RTGS_RENDER_MODE=0
run slot_mode2 RTGS_B200_LIB=$L/librtgs_slot.so
done
cat gpurun_out/r2_ab1.log
