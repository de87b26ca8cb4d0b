#!/bin/bash
# usage: scratch/final_profile.sh TAG -> GPU tests, smoke, default bench (+ reference arm), launch list, ncu --set full
tag=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1
python bench.py > gpurun_out/bench_$tag.log 2>&1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$tag.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_tile_lists|k_shade_tiles' --launch-skip 8 --launch-count 2 \
    -f -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
