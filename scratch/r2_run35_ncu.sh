#!/bin/bash
export RTGS_HEAVY_SLAB=2
B="python bench.py --config surface_1m_1080p --steps 2 --warmup 2 --no-cpu-baseline --no-tiles"
$B > gpurun_out/r02_heavy_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_heavy_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_heavy_lists --launch-skip 4 --launch-count 1 -f -o gpurun_out/r02_heavy $B > gpurun_out/r02_ncu_heavy.log 2>&1
ls -la gpurun_out/r02_heavy.ncu-rep; tail -2 gpurun_out/r02_ncu_heavy.log
