#!/bin/bash
for sm in 16 24; do
echo "== RTGS_SLAB_RANK=$sm"
RTGS_SLAB_RANK=$sm timeout 600 python scratch/heavy_probe.py 2>&1 | grep "heavy_lists 2"
done > gpurun_out/r2_heavy_probe32.log 2>&1
cat gpurun_out/r2_heavy_probe32.log
