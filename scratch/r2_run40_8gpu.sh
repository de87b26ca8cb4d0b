#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29721 bench.py --gpus 8 --steps 64 --warmup 5 > gpurun_out/r2_bench40_n8.log 2> gpurun_out/r2_bench40_n8.err
echo "bench n8 rc=$?"; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/r2_bench40_n8.err | tail -5
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench40_n8.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'e2e',d['e2e']['value'])
        for key in ('tiles','tiles_config4'):
            t=d[key]; print(key,{k:v for k,v in t.items() if k not in ('by_mode','single_gpu_by_mode','limiting_kernel','e2e')})
            for m,v in t['by_mode'].items(): print('  ',m, round(v['ms_per_frame_latency'],4), round(v['ms_per_frame_two_streams'],4), [round(max(r[i] for r in v['kernels_ms_per_rank']),4) for i in range(3)])
PY
