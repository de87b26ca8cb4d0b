#!/bin/bash
# 2 GPUs: the default bench line (views gathered on GPU 0 + tile-sharded frames), the D2H probe at 1 and 2 processes
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29501 bench.py --gpus 2 --steps 64 --warmup 5 > gpurun_out/r2_bench7_n2.log 2> gpurun_out/r2_bench7_n2.err
echo "bench rc=$?"; tail -c 4500 gpurun_out/r2_bench7_n2.log; tail -5 gpurun_out/r2_bench7_n2.err
timeout 120 python tools/d2h_probe.py > gpurun_out/r2_d2h_n1.log 2>&1; cat gpurun_out/r2_d2h_n1.log | tail -2
timeout 120 $TR --nproc-per-node 2 --master-port 29502 tools/d2h_probe.py > gpurun_out/r2_d2h_n2.log 2>&1; tail -2 gpurun_out/r2_d2h_n2.log
timeout 300 $TR --nproc-per-node 2 --master-port 29503 bench.py --gpus 2 --steps 20 --warmup 3 --impl reference > gpurun_out/r2_ref7_n2.log 2>&1; tail -c 600 gpurun_out/r2_ref7_n2.log
