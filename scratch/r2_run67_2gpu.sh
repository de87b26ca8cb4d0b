#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29731 bench.py --gpus 2 --steps 48 --warmup 5 > gpurun_out/r2_bench67_n2.log 2> gpurun_out/r2_bench67_n2.err
echo "bench rc=$?"; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/r2_bench67_n2.err | tail -5
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench67_n2.log'):
    if l.startswith('{'):
        d=json.loads(l); t=d['tiles']
        print('value',d['value'],'e2e',d['e2e']['value']); print({k:v for k,v in t.items() if k not in ('by_mode','single_gpu_by_mode','limiting_kernel','e2e')})
        for m,v in t['by_mode'].items(): print(m, v['ms_per_frame_latency'], v['ms_per_frame_two_streams'])
PY
