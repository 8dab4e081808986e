#!/bin/bash
# Second front-end call: the WavLM bf16 mode (tensor-core attention with the position bias), the faster positional convolution, and the
# whole GPU suite again (flash kernel / shared encoder header touched).
mkdir -p gpurun_out/r2_frontend2
timeout 200 python -m pytest tests/test_zz_wavlm_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_frontend2/wavlm_tests.log 2>&1
tail -30 gpurun_out/r2_frontend2/wavlm_tests.log
timeout 120 python tools/measure_frontend.py 6 > gpurun_out/r2_frontend2/frontend.log 2>&1
tail -5 gpurun_out/r2_frontend2/frontend.log
timeout 300 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/r2_frontend2/gpu_suite.log 2>&1
tail -8 gpurun_out/r2_frontend2/gpu_suite.log
