#!/bin/bash
timeout 300 python scratch/dbg_fuzz2_px.py 25 19 24 > gpurun_out/r2_dbg_px.log 2>&1; cat gpurun_out/r2_dbg_px.log | tail -40
