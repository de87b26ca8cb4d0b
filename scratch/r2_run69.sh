#!/bin/bash
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest69.log 2>&1; tail -3 gpurun_out/r2_pytest69.log
