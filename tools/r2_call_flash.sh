#!/bin/bash
# flash attention + fused bf16 block path: parity tests, kernel timing, estimator call timing (each step under its own short timeout)
mkdir -p gpurun_out/r2_fl
timeout 120 python -m pytest tests/test_flash_gpu.py -x -q 2>&1 | tail -3 || exit 1
timeout 300 python -m pytest tests/test_zz_options_gpu.py tests/test_zz_dit_gpu.py -x -q -k "dit" 2>&1 | tail -5
timeout 120 python tools/measure_dit.py --bf16 --reps 10 2>&1 | tail -6 | tee gpurun_out/r2_fl/dit.log
