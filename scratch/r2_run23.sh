#!/bin/bash
for lim in 0 72 96 128 192; do
echo "== RTGS_FIRST_LIMIT=$lim" >> gpurun_out/r2_first23.log
RTGS_FIRST_LIMIT=$lim python scratch/stripe_probe.py 1m_deg3_1080p 0 2>&1 | grep "world 1 \|world 8 rank 0" >> gpurun_out/r2_first23.log
done
for lim in 0 96 160; do
echo "== 4K RTGS_FIRST_LIMIT=$lim" >> gpurun_out/r2_first23.log
RTGS_FIRST_LIMIT=$lim python scratch/stripe_probe.py 3m_deg3_2160p 0 2>&1 | grep "world 1 \|world 8 rank 0" >> gpurun_out/r2_first23.log
done
cat gpurun_out/r2_first23.log
RTGS_FIRST_LIMIT=96 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest23.log 2>&1; tail -3 gpurun_out/r2_pytest23.log
