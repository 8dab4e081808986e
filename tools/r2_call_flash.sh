#!/bin/bash
# flash attention: parity tests, kernel timing, estimator call timing (every step under its own short timeout: a barrier bug would hang)
mkdir -p gpurun_out/r2_fl
timeout 60 python -m pytest tests/test_flash_gpu.py -x -q -k "1-2-129" 2>&1 | tail -5 || exit 1
timeout 120 python -m pytest tests/test_flash_gpu.py -x -q 2>&1 | tail -15 || exit 1
timeout 60 python tools/measure_flash.py 2>&1 | tail -8 | tee gpurun_out/r2_fl/flash.log
timeout 120 python -m pytest tests/test_zz_options_gpu.py -x -q -k "dit_bf16" 2>&1 | tail -5
timeout 120 python tools/measure_dit.py --bf16 --reps 10 2>&1 | tail -6 | tee gpurun_out/r2_fl/dit.log
