#!/bin/bash
timeout 900 python scratch/fuzz4.py 0 200 > gpurun_out/r2_fz4.log 2>&1; tail -2 gpurun_out/r2_fz4.log | cut -c1-200; grep "FAIL\|Error\|error" gpurun_out/r2_fz4.log | head -8 | cut -c1-260; true
