#!/bin/bash
# round 2, GPU run 2: where did the regression come from (rdc? ldcg?), and does the one-launch frame win at small shares?
L=$PWD/rt-gaussian-splat-renderer_b200/lib
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_bench_$tag.err | python -c "
import sys,json
for x in sys.stdin:
    if x.startswith('{'):
        d=json.loads(x); print('$tag', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'sync', round(d['e2e']['sync_value'],1), 'ms', round(d['ms_per_step'],4), [(k['kernel'], round(k['ms'],4)) for k in d['kernels']])
" >> gpurun_out/r2_ab2.log
}
for i in 1 2; do
run t16_m0 RTGS_B200_LIB=$L/lib_t16.so RTGS_RENDER_MODE=0
run t16_ldg_m0 RTGS_B200_LIB=$L/lib_t16_ldg.so RTGS_RENDER_MODE=0
run t16_cdp_m0 RTGS_B200_LIB=$L/lib_t16_cdp.so RTGS_RENDER_MODE=0
run t16_m2 RTGS_B200_LIB=$L/lib_t16.so RTGS_RENDER_MODE=2
done
cat gpurun_out/r2_ab2.log
RTGS_B200_LIB=$L/lib_t16.so python scratch/stripe_probe.py 1m_deg3_1080p 0 2 > gpurun_out/r2_stripe_t16.log 2>&1
RTGS_B200_LIB=$L/lib_t32.so python scratch/stripe_probe.py 1m_deg3_1080p 0 2 > gpurun_out/r2_stripe_t32.log 2>&1
cat gpurun_out/r2_stripe_t16.log gpurun_out/r2_stripe_t32.log
