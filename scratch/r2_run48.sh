#!/bin/bash
timeout 1700 python scratch/fuzz2.py 0 200 > gpurun_out/r2_fuzz2.log 2>&1; tail -2 gpurun_out/r2_fuzz2.log; grep FAIL gpurun_out/r2_fuzz2.log | head -20; true
