#!/bin/bash
timeout 1500 python scratch/fuzz_heavy.py 0 150 > gpurun_out/r2_fuzz_heavy.log 2>&1; tail -4 gpurun_out/r2_fuzz_heavy.log; grep -c FAIL gpurun_out/r2_fuzz_heavy.log
