#!/bin/bash
timeout 600 python scratch/dbg_seed20b.py > gpurun_out/r2_dbg_seed20.log 2>&1; cat gpurun_out/r2_dbg_seed20.log | tail -30
