#!/bin/bash
mkdir -p gpurun_out
python tools/mm_probe.py 1024 > gpurun_out/s9_mm_probe.txt 2>&1; cat gpurun_out/s9_mm_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mml -c 600 --csv --log-file gpurun_out/s9_mm_launches.csv python tools/mm_probe.py 1024 > gpurun_out/s9_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/s9_mm_launches.csv | head -20
