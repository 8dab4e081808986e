#!/bin/bash
# ncu launch list + one full capture of the new front-end kernels (WavLM base-plus forward + get_whisper_features, 6 x 30 s)
mkdir -p gpurun_out/r2_frontend
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_frontend/launches.csv \
    python tools/profile_frontend.py 6 > gpurun_out/r2_frontend/prof1.log 2>&1
tail -2 gpurun_out/r2_frontend/prof1.log; wc -l gpurun_out/r2_frontend/launches.csv
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"fe_logmel_kernel|wl_posconv_kernel|dit_attn_kernel|wl_conv0_apply_kernel" -c 4 -o gpurun_out/r2_frontend/prof \
    python tools/profile_frontend.py 6 > gpurun_out/r2_frontend/prof2.log 2>&1
tail -3 gpurun_out/r2_frontend/prof2.log; ls -la gpurun_out/r2_frontend/
