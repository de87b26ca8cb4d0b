#!/bin/bash
# 8 GPUs: the default bench line as the driver runs it (views gathered on GPU 0, tile-sharded config 3 and config 4),
# the reference arm's core count, and the D2H ceiling at 4 and 8 processes
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo8.log 2>&1
timeout 900 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 64 --warmup 5 > gpurun_out/r2_bench11_n8.log 2> gpurun_out/r2_bench11_n8.err
echo "bench n8 rc=$?"; tail -c 1500 gpurun_out/r2_bench11_n8.log; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/r2_bench11_n8.err | tail -8
for n in 8 4 2 1; do
  timeout 120 $TR --nproc-per-node $n --master-port 2952$n tools/d2h_probe.py > gpurun_out/r2_d2h11_n$n.log 2>&1; grep '^{' gpurun_out/r2_d2h11_n$n.log
done
timeout 200 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --impl reference > gpurun_out/r2_ref11_n8.log 2>&1; grep -o '"cores": [0-9]*' gpurun_out/r2_ref11_n8.log | head -2
