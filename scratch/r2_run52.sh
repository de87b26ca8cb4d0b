#!/bin/bash
RTGS_HEAVY_FUSED=0 timeout 300 python scratch/dbg_fuzz2.py 25 A 2>&1 | grep "mode" | cut -c1-260 > gpurun_out/r2_dbg52.log; cat gpurun_out/r2_dbg52.log
