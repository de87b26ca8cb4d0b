#!/bin/bash
timeout 300 python scratch/dbg_fuzz2_region.py > gpurun_out/r2_dbg53.log 2>&1; tail -12 gpurun_out/r2_dbg53.log | cut -c1-330
