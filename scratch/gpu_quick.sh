#!/bin/bash
# usage: scratch/gpu_quick.sh tag  -> GPU tests + bench + per-kernel launch list
tag=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1
python bench.py --steps 64 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
