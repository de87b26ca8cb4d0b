#!/bin/bash
timeout 1500 python scratch/fullframe_check.py 1m_deg3_1080p 7 40 > gpurun_out/r2_fullframe39_cfg3.log 2>&1; tail -5 gpurun_out/r2_fullframe39_cfg3.log
