#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest27.log 2>&1; tail -3 gpurun_out/r2_pytest27.log
timeout 900 python scratch/fullframe_check.py 1m_deg3_1080p 0 21 > gpurun_out/r2_fullframe27_cfg3.log 2>&1; cat gpurun_out/r2_fullframe27_cfg3.log | grep -v worst
timeout 900 python scratch/fullframe_surface.py 0 5 > gpurun_out/r2_fullframe27_surface.log 2>&1; cat gpurun_out/r2_fullframe27_surface.log
timeout 900 python scratch/fullframe_check.py 3m_deg3_2160p 7 > gpurun_out/r2_fullframe27_cfg4.log 2>&1; cat gpurun_out/r2_fullframe27_cfg4.log | grep -v worst
