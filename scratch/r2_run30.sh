#!/bin/bash
timeout 600 python scratch/heavy_probe.py > gpurun_out/r2_heavy_probe.log 2>&1; cat gpurun_out/r2_heavy_probe.log | tail -20
