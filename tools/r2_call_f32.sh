#!/bin/bash
mkdir -p gpurun_out/r2_f32
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_f32/frame32_launches.csv python tools/profile_frame.py --batch 32 --frames 2 > gpurun_out/r2_f32/prof.log 2>&1
tail -3 gpurun_out/r2_f32/prof.log
