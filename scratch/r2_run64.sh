#!/bin/bash
python bench.py > gpurun_out/r2_bench64.json 2> gpurun_out/r2_bench64.err; tail -2 gpurun_out/r2_bench64.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench64.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'serial',round(d['serial_value'],1),'e2e',round(d['e2e']['value'],1),'sync',round(d['e2e']['sync_value'],1),'compact',round(d['e2e']['compact']['value'],1),'roofline',round(d['roofline']['frac'],3),[(k['kernel'],round(k['ms'],4)) for k in d['kernels']],'cpu',d['cpu_baseline']['value'],'launches',d['gpu_launches'],'clocks',d['clocks']['sm_mhz'],d['clocks']['reasons'])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
