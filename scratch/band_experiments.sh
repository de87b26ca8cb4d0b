#!/bin/bash
# the cost of the banded host delivery, one knob at a time (each run: device-timed value, e2e, kernel times in e2e mode)
run() { tag=$1; shift; env "$@" python bench.py --steps 64 --warmup 5 --no-cpu-baseline > gpurun_out/band_$tag.log 2>&1; }
run base X=1
run sched0 RTGS_BAND_SCHEDULE=0
run nosignal RTGS_BAND_NOSIGNAL=1
run nocopy RTGS_BAND_NOCOPY=1
