#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 64 --warmup 5 > gpurun_out/r2_bench22_n8.log 2> gpurun_out/r2_bench22_n8.err
echo "bench n8 rc=$?"; tail -c 600 gpurun_out/r2_bench22_n8.log; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/r2_bench22_n8.err | tail -5
timeout 600 $TR --nproc-per-node 4 --master-port 29612 bench.py --gpus 4 --steps 64 --warmup 5 > gpurun_out/r2_bench22_n4.log 2> gpurun_out/r2_bench22_n4.err
echo "bench n4 rc=$?"; tail -c 300 gpurun_out/r2_bench22_n4.log
