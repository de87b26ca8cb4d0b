#!/bin/bash
# ncu inventory of round 2 (B200_PROFILING.md recipe): launch list, then --set full per kernel family
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-tiles"
$B > gpurun_out/r02_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_tile_lists|k_shade_tiles' --launch-skip 8 --launch-count 2 -f -o gpurun_out/r02_render $B > gpurun_out/r02_ncu_render.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_pack|k_refit|k_rs_scatter|k_rs_hist|k_karras|k_morton|k_bounds|k_max_depth' --launch-count 14 -f -o gpurun_out/r02_build $B > gpurun_out/r02_ncu_build.log 2>&1
RTGS_RENDER_MODE=2 ncu --set full --clock-control none --import-source on -k regex:'k_frame' --launch-skip 4 --launch-count 1 -f -o gpurun_out/r02_frame $B > gpurun_out/r02_ncu_frame.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render --launch-skip 6 --launch-count 1 -f -o gpurun_out/r02_krender_surface python bench.py --config surface_1m_1080p --steps 2 --warmup 2 --no-cpu-baseline --no-tiles > gpurun_out/r02_ncu_krender.log 2>&1
python scratch/trace_probe.py > gpurun_out/r02_trace_plain.log 2>&1; tail -1 gpurun_out/r02_trace_plain.log
ncu --set full --clock-control none --import-source on -k regex:k_trace_closest --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_trace python scratch/trace_probe.py > gpurun_out/r02_ncu_trace.log 2>&1
ls -la gpurun_out/r02_*.ncu-rep
# memcheck of the deep-tree tests
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck_deep_tree.log python -m pytest tests/test_gpu_deep_tree.py -x -q > gpurun_out/r02_memcheck_pytest.log 2>&1; tail -3 gpurun_out/r02_memcheck_pytest.log; tail -5 gpurun_out/r02_memcheck_deep_tree.log
