#!/bin/bash
mkdir -p gpurun_out/r2_pw
timeout 300 python -m pytest tests/test_zz_options_gpu.py tests/test_codec_gpu.py tests/test_scalar_gpu.py -x -q -k "pointwise or codec or conv or scalar" 2>&1 | tail -4
timeout 200 python tools/measure_codec.py conv_pointwise 2>&1 | tail -3 | tee gpurun_out/r2_pw/codec.log
timeout 200 python tools/measure_kernels.py 2>&1 | grep "enc res\|dec up k8" | cut -c1-330 | tee gpurun_out/r2_pw/kernels.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_pw/codec_launches.csv python tools/profile_codec.py --batch 16 --seconds 10 --reps 1 > gpurun_out/r2_pw/prof.log 2>&1
