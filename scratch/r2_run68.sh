#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r2_pytest68.log 2>&1; tail -12 gpurun_out/r2_pytest68.log | cut -c1-250
