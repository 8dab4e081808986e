#!/bin/bash
# Round 2: bench (default) + codec sweep + frame launch list + batch-32 frame measurement.  Everything lands in gpurun_out/r2_f1/.
set -u
out=gpurun_out/r2_f1
mkdir -p "$out"
python -c "import torch" > /dev/null 2>&1
timeout -k 5 600 python bench.py --steps 3 --warmup 3 --codec-sweep > "$out/bench.json" 2> "$out/bench.err"; tail -c 1500 "$out/bench.json"; tail -3 "$out/bench.err"
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/frame_launches.csv" python tools/profile_frame.py --frames 3 > "$out/frame_launches.log" 2>&1; tail -2 "$out/frame_launches.log"
timeout -k 5 300 python tools/measure_configs.py --only prefill,caption32,ttm500 > "$out/configs.log" 2>&1; tail -12 "$out/configs.log" | cut -c1-300
timeout -k 5 200 python tools/measure_kernels.py --conv-tc --resblock --attn-ring > "$out/kernels.log" 2>&1; tail -3 "$out/kernels.log" | cut -c1-200
