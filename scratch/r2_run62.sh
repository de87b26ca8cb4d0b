#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "as_given" 2>&1 | tail -4
timeout 600 python scratch/fuzz4.py 55 62 2>&1 | tail -4 | cut -c1-200
