#!/bin/bash
mkdir -p gpurun_out/r2_sq
timeout 300 python -m pytest tests/test_zz_options_gpu.py tests/test_scalar_gpu.py tests/test_codec_gpu.py -q -k "scalar or conv or codec" 2>&1 | tail -8
timeout 200 python tools/measure_scalar.py 1 2>&1 | tail -1 | tee gpurun_out/r2_sq/scalar.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_sq/scalar_launches.csv python tools/measure_scalar.py 1 --once > gpurun_out/r2_sq/prof.log 2>&1
