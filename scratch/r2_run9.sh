#!/bin/bash
timeout 600 python scratch/surface_stats.py > gpurun_out/r2_surface_stats.log 2>&1; cat gpurun_out/r2_surface_stats.log | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render --launch-skip 6 --launch-count 1 -o gpurun_out/r02_krender_surface python bench.py --config surface_1m_1080p --steps 2 --warmup 2 --no-cpu-baseline --no-tiles > gpurun_out/r2_ncu9.log 2>&1; tail -3 gpurun_out/r2_ncu9.log; ls -la gpurun_out/*.ncu-rep
