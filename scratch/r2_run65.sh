#!/bin/bash
timeout 600 python scratch/fuzz5.py 0 150 > gpurun_out/r2_fz5.log 2>&1; tail -3 gpurun_out/r2_fz5.log | cut -c1-200; grep "FAIL\|Error" gpurun_out/r2_fz5.log | head -6 | cut -c1-250; true
