#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest59.log 2>&1; tail -2 gpurun_out/r2_pytest59.log
( timeout 700 python scratch/fuzz.py 120 400 > gpurun_out/r2_fz_a.log 2>&1; tail -1 gpurun_out/r2_fz_a.log; grep -c FAIL gpurun_out/r2_fz_a.log ) 
( timeout 900 python scratch/fuzz2.py 330 700 > gpurun_out/r2_fz_b.log 2>&1; tail -1 gpurun_out/r2_fz_b.log; grep FAIL gpurun_out/r2_fz_b.log | head -5 | cut -c1-300 )
( timeout 900 python scratch/fuzz3.py 100 260 > gpurun_out/r2_fz_c.log 2>&1; tail -1 gpurun_out/r2_fz_c.log; grep FAIL gpurun_out/r2_fz_c.log | head -5 | cut -c1-300 )
( timeout 600 python scratch/fuzz_heavy.py 330 480 > gpurun_out/r2_fz_d.log 2>&1; tail -1 gpurun_out/r2_fz_d.log; grep FAIL gpurun_out/r2_fz_d.log | head -5 | cut -c1-300 )
true
