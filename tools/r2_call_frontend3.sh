#!/bin/bash
# Third front-end call: log-mel v2 (twiddle rotation in registers), timings, ncu launch list of the bf16-mode step.
mkdir -p gpurun_out/r2_frontend3
timeout 200 python -m pytest tests/test_zz_frontend_gpu.py tests/test_zz_wavlm_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_frontend3/tests.log 2>&1
tail -20 gpurun_out/r2_frontend3/tests.log
timeout 120 python tools/measure_frontend.py 6 > gpurun_out/r2_frontend3/frontend.log 2>&1
tail -5 gpurun_out/r2_frontend3/frontend.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_frontend3/launches_bf16.csv \
    python tools/profile_frontend.py 6 bf16 > gpurun_out/r2_frontend3/prof1.log 2>&1
wc -l gpurun_out/r2_frontend3/launches_bf16.csv
