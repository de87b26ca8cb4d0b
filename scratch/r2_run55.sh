#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest55.log 2>&1; tail -3 gpurun_out/r2_pytest55.log
timeout 1500 python scratch/fuzz3.py 0 100 > gpurun_out/r2_fuzz3.log 2>&1; tail -3 gpurun_out/r2_fuzz3.log | cut -c1-300; grep FAIL gpurun_out/r2_fuzz3.log | head -5 | cut -c1-300; true
