#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke36.log 2>&1; tail -2 gpurun_out/r2_smoke36.log
( time python bench.py ) > gpurun_out/r2_bench36.json 2> gpurun_out/r2_bench36.err; tail -4 gpurun_out/r2_bench36.err; head -c 3000 gpurun_out/r2_bench36.json
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench36_ref.json 2> gpurun_out/r2_bench36_ref.err; tail -4 gpurun_out/r2_bench36_ref.err; cat gpurun_out/r2_bench36_ref.json
