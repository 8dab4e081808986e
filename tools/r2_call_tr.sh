#!/bin/bash
# truncating on-chip split: parity first (full GPU suite), then the numbers it should move
mkdir -p gpurun_out/r2_tr
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_tr/tests.log
timeout 100 python tools/check_umma.py 2>&1 | grep '"M"' | cut -c1-190 | tee gpurun_out/r2_tr/umma.log
timeout 100 python tools/measure_codec.py conv_umma 2>&1 | tail -1 | tee gpurun_out/r2_tr/codec.log
timeout 100 python tools/measure_scalar.py 1 2>&1 | tail -1 | tee gpurun_out/r2_tr/scalar.log
timeout 200 python tools/measure_configs.py --only prefill,caption32 2>&1 | grep tcgen05 | cut -c1-330 | tee gpurun_out/r2_tr/configs.log
