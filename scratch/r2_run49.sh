#!/bin/bash
for a in "25 A" "57 B" "91 C"; do echo "=== $a"; timeout 300 python scratch/dbg_fuzz2.py $a 2>&1 | tail -14 | cut -c1-400; done > gpurun_out/r2_dbg_fuzz2.log 2>&1
cat gpurun_out/r2_dbg_fuzz2.log
