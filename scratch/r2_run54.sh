#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -x -q -s -k "dense_cluster" > gpurun_out/r2_pytest54a.log 2>&1; tail -6 gpurun_out/r2_pytest54a.log
for a in "25 A" "57 B" "91 C"; do echo "=== $a"; timeout 300 python scratch/dbg_fuzz2.py $a 2>&1 | grep "mode 0\|mode 2\|mode 1" | head -5 | cut -c1-200; done
timeout 1500 python scratch/fuzz2.py 0 200 > gpurun_out/r2_fuzz2b.log 2>&1; tail -2 gpurun_out/r2_fuzz2b.log; grep FAIL gpurun_out/r2_fuzz2b.log | head; true
