#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest63.log 2>&1; tail -2 gpurun_out/r2_pytest63.log
timeout 600 python scratch/fuzz4.py 200 420 > gpurun_out/r2_fz4b.log 2>&1; tail -1 gpurun_out/r2_fz4b.log; grep "FAIL\|Error" gpurun_out/r2_fz4b.log | head -5 | cut -c1-250; true
