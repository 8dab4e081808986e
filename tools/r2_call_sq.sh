#!/bin/bash
mkdir -p gpurun_out/r2_sq
timeout 300 python -m pytest tests/test_zz_options_gpu.py tests/test_scalar_gpu.py tests/test_codec_gpu.py -q -k "scalar or conv or codec" 2>&1 | tail -4
timeout 200 python tools/measure_scalar.py 1 2>&1 | tail -1 | tee gpurun_out/r2_sq/scalar.log
timeout 200 python tools/measure_codec.py conv_umma_staged 2>&1 | tail -2 | tee gpurun_out/r2_sq/codec.log
timeout 200 python tools/measure_kernels.py 2>&1 | grep "enc \|dec up" | cut -c1-200 | tee gpurun_out/r2_sq/kernels.log
