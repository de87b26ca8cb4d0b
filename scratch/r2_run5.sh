#!/bin/bash
L=$PWD/rt-gaussian-splat-renderer_b200/lib
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "sync", round(d["e2e"]["sync_value"],1), "ms", round(d["ms_per_step"],4), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-tiles 2> gpurun_out/r2_bench_$tag.err | python -c "$fmt" $tag >> gpurun_out/r2_ab5.log
}
run m0_single112 RTGS_RENDER_MODE=0
run m0_single64 RTGS_RENDER_MODE=0 RTGS_LISTS_SINGLE=64
run m0_single88 RTGS_RENDER_MODE=0 RTGS_LISTS_SINGLE=88
run m0_single140 RTGS_RENDER_MODE=0 RTGS_LISTS_SINGLE=140
run m0_single112_fb1 RTGS_RENDER_MODE=0 RTGS_FB_GRID=1
run m0_s512_single112 RTGS_RENDER_MODE=0 RTGS_B200_LIB=$L/lib_s512.so
run m0_s512_single320 RTGS_RENDER_MODE=0 RTGS_B200_LIB=$L/lib_s512.so RTGS_LISTS_SINGLE=320
cat gpurun_out/r2_ab5.log
python scratch/stripe_probe.py 1m_deg3_1080p 0 2 > gpurun_out/r2_stripe5_t16.log 2>&1
RTGS_B200_LIB=$L/lib_t32.so python scratch/stripe_probe.py 1m_deg3_1080p 2 > gpurun_out/r2_stripe5_t32.log 2>&1
cat gpurun_out/r2_stripe5_t16.log gpurun_out/r2_stripe5_t32.log
# the full bench line once (tiles section on one GPU: trivially world = 1)
timeout 600 python bench.py --steps 64 --warmup 5 > gpurun_out/r2_bench5_full.log 2> gpurun_out/r2_bench5_full.err; tail -c 3000 gpurun_out/r2_bench5_full.log; tail -5 gpurun_out/r2_bench5_full.err
