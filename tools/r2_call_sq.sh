#!/bin/bash
mkdir -p gpurun_out/r2_sq
timeout 200 python tools/measure_codec.py conv_umma_k1 2>&1 | tail -2 | tee gpurun_out/r2_sq/codec_k1.log
UA2_OPTIONS="conv_umma_k1=1" timeout 200 python tools/measure_kernels.py 2>&1 | grep "enc res k1" | cut -c1-200 | tee gpurun_out/r2_sq/kernels_k1.log
UA2_OPTIONS="conv_umma_k1=1" timeout 200 python -m pytest tests/test_codec_gpu.py -q 2>&1 | tail -2
UA2_OPTIONS="conv_umma_k1=1" timeout 100 python tools/measure_scalar.py 1 2>&1 | tail -1
