#!/bin/bash
timeout 1500 python scratch/fullframe_clusters.py > gpurun_out/r2_fullframe_clusters.log 2>&1; tail -16 gpurun_out/r2_fullframe_clusters.log | cut -c1-300

