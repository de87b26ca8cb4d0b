#!/bin/bash
timeout 600 python scratch/heavy_probe.py > gpurun_out/r2_heavy_probe.log 2>&1; cat gpurun_out/r2_heavy_probe.log | tail -20
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -x -q -s -k "heavy or overflowing or render_paths" > gpurun_out/r2_pytest31.log 2>&1; tail -8 gpurun_out/r2_pytest31.log
