#!/bin/bash
# Final GPU call of round 2: the new tokenize front-ends first (their first hardware run), then the whole GPU suite (regression for the
# shared encoder header / the attention kernel's new template parameter), then timings and the smoke entry.
mkdir -p gpurun_out/r2_frontend
timeout 240 python -m pytest tests/test_zz_frontend_gpu.py tests/test_zz_wavlm_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_frontend/new_tests.log 2>&1
tail -25 gpurun_out/r2_frontend/new_tests.log
timeout 120 python tools/measure_frontend.py 6 > gpurun_out/r2_frontend/frontend.log 2>&1
tail -4 gpurun_out/r2_frontend/frontend.log
timeout 300 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/r2_frontend/gpu_suite.log 2>&1
tail -6 gpurun_out/r2_frontend/gpu_suite.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/r2_frontend/smoke.log
