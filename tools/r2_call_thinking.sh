#!/bin/bash
# First hardware run of the reasoning encoder (AudioThinking) + timing at the reference's geometry
mkdir -p gpurun_out/r2_thinking
timeout 200 python -m pytest tests/test_zz_thinking_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_thinking/tests.log 2>&1
tail -25 gpurun_out/r2_thinking/tests.log
timeout 100 python tools/measure_thinking.py 6 > gpurun_out/r2_thinking/thinking.log 2>&1
tail -3 gpurun_out/r2_thinking/thinking.log
