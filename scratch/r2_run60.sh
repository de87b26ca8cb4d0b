#!/bin/bash
timeout 900 python scratch/fuzz_stripes.py 0 120 > gpurun_out/r2_fz_stripes.log 2>&1; tail -2 gpurun_out/r2_fz_stripes.log | cut -c1-200; grep FAIL gpurun_out/r2_fz_stripes.log | head -5 | cut -c1-200; true
