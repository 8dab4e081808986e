#!/bin/bash
mkdir -p gpurun_out/r2_prof
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 1 -c 1 -o gpurun_out/r2_prof/convumma_staged python tools/profile_convumma.py > gpurun_out/r2_prof/ncu.log 2>&1
tail -3 gpurun_out/r2_prof/ncu.log
ls -la gpurun_out/r2_prof/
