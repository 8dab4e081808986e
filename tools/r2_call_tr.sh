#!/bin/bash
# sleeping waits in the non-critical warps of the tcgen05 kernels: parity (full GPU suite), then the numbers it should move
mkdir -p gpurun_out/r2_tr
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_tr/tests.log
timeout 100 python tools/check_umma.py 2>&1 | grep 'weight_GBps' | cut -c1-200 | tee gpurun_out/r2_tr/umma.log
timeout 100 python tools/measure_codec.py conv_umma 2>&1 | tail -1 | tee gpurun_out/r2_tr/codec.log
timeout 100 python tools/measure_scalar.py 1 2>&1 | tail -1 | tee gpurun_out/r2_tr/scalar.log
timeout 100 python tools/measure_flash.py 2>&1 | grep "<1>" | cut -c1-120 | tee gpurun_out/r2_tr/flash.log
timeout 100 python tools/measure_dit.py --bf16 --reps 5 2>&1 | grep "tensor-core\|3xTF32" | cut -c1-200 | tee gpurun_out/r2_tr/dit.log
timeout 200 python tools/measure_configs.py --only prefill,caption32 2>&1 | grep tcgen05 | cut -c1-330 | tee gpurun_out/r2_tr/configs.log
