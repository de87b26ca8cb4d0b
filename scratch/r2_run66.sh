#!/bin/bash
timeout 700 python scratch/fuzz6.py 0 120 > gpurun_out/r2_fz6.log 2>&1; tail -2 gpurun_out/r2_fz6.log | cut -c1-250; grep "FAIL\|Error" gpurun_out/r2_fz6.log | head -6 | cut -c1-300; true
