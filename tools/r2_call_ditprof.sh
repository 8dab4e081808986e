#!/bin/bash
mkdir -p gpurun_out/r2_dp
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_dp/dit_bf16_launches.csv python tools/measure_dit.py --once --bf16 > gpurun_out/r2_dp/once.log 2>&1
tail -2 gpurun_out/r2_dp/once.log
