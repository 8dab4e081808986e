#!/bin/bash
mkdir -p gpurun_out/r2_dbg
timeout 600 compute-sanitizer --tool initcheck --print-limit 30 python -m pytest tests/test_zz_whisper_gpu.py -x -q -k "128-2-512" > gpurun_out/r2_dbg/initcheck.log 2>&1
grep -c "Uninitialized" gpurun_out/r2_dbg/initcheck.log
grep -A12 "Uninitialized" gpurun_out/r2_dbg/initcheck.log | head -80
tail -5 gpurun_out/r2_dbg/initcheck.log
timeout 300 python -m pytest tests/test_zz_options_gpu.py tests/test_zz_whisper_gpu.py -x -q 2>&1 | tail -5
