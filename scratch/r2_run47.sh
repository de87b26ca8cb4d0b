#!/bin/bash
timeout 1500 python scratch/fuzz_heavy.py 150 330 > gpurun_out/r2_fuzz_heavy3.log 2>&1; tail -2 gpurun_out/r2_fuzz_heavy3.log; grep -c FAIL gpurun_out/r2_fuzz_heavy3.log
timeout 1200 python scratch/fuzz.py 0 120 > gpurun_out/r2_fuzz47.log 2>&1; tail -1 gpurun_out/r2_fuzz47.log; grep -c FAIL gpurun_out/r2_fuzz47.log
