#!/bin/bash
L=$PWD/rt-gaussian-splat-renderer_b200/lib
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1; tail -3 gpurun_out/r2_pytest4.log
fmt='
import sys,json
for x in sys.stdin:
    if x.startswith("{"):
        d=json.loads(x); print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "sync", round(d["e2e"]["sync_value"],1), "ms", round(d["ms_per_step"],4), [(k["kernel"], round(k["ms"],4)) for k in d["kernels"]])
'
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2_bench_$tag.err | python -c "$fmt" $tag >> gpurun_out/r2_ab4.log
}
(cd rt-gaussian-splat-renderer_b200/build/r1 && timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2> /dev/null | python -c "$fmt" r1_baseline >> ../../../gpurun_out/r2_ab4.log)
run cur_m0 RTGS_RENDER_MODE=0
run cur_m0_1stream RTGS_RENDER_MODE=0 RTGS_SUBMIT_ONE_STREAM=1
run v4_m0 RTGS_B200_LIB=$L/lib_v4.so RTGS_RENDER_MODE=0
run v5_m0 RTGS_B200_LIB=$L/lib_v5.so RTGS_RENDER_MODE=0
run cur_m2 RTGS_RENDER_MODE=2
(cd rt-gaussian-splat-renderer_b200/build/r1 && timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2> /dev/null | python -c "$fmt" r1_baseline >> ../../../gpurun_out/r2_ab4.log)
cat gpurun_out/r2_ab4.log
