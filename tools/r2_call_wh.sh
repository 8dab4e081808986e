#!/bin/bash
mkdir -p gpurun_out/r2_wh
timeout 300 python -m pytest tests/test_zz_whisper_gpu.py -x -q 2>&1 | tail -15
timeout 200 python tools/measure_whisper.py 6 2>&1 | tail -4 | tee gpurun_out/r2_wh/whisper.log
