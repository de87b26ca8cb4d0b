#!/bin/bash
for lib in librtgs_old.so librtgs_b200.so; do
for a in "25 A" "91 C"; do echo "=== $lib $a"; RTGS_B200_LIB=$PWD/rt-gaussian-splat-renderer_b200/lib/$lib timeout 300 python scratch/dbg_fuzz2.py $a 2>&1 | grep -v "^{" | grep "mode 0\|mode 1" | head -6 | cut -c1-300; done; done > gpurun_out/r2_dbg_fuzz2b.log 2>&1
cat gpurun_out/r2_dbg_fuzz2b.log
