#!/bin/bash
tag=$1
tail -2 gpurun_out/pytest_$tag.log
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/bench_$tag.log") if x.startswith("{")][-1]); print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],4))
PY
grep -v "^==" gpurun_out/launches_$tag.csv | awk -F'","' '{print $5, $NF}' | tail -6 | sed 's/void <unnamed>:://; s/(rtgs_dev::RenderParams)//'
