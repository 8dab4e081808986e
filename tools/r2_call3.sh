#!/bin/bash
# Round 2, call 3: tests on the elect-issued tcgen05 kernel, codec launch list, ncu captures of the attention and SEANet kernels.
set -u
out=gpurun_out/r2_c3
mkdir -p "$out"
python -c "import torch" > /dev/null 2>&1
timeout -k 5 600 python -m pytest tests -q -m gpu -p no:cacheprovider > "$out/tests.log" 2>&1; tail -5 "$out/tests.log"
timeout -k 5 120 python tools/check_umma.py --time > "$out/check_umma.log" 2>&1; grep -c '"ok": true' "$out/check_umma.log"; grep weight_GBps "$out/check_umma.log" | cut -c1-150 | sed -n '10,13p'
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/codec_launches.csv" python tools/profile_codec.py --batch 16 --seconds 10 --reps 2 > "$out/codec_launches.log" 2>&1; tail -2 "$out/codec_launches.log"
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:"attn_split_kernel|attn_ring_kernel|resblock64_kernel|sgemm_conv_kernel|umma_kernel" -o "$out/attn_conv" -f python tools/profile_attn_conv.py > "$out/ncu_attn_conv.log" 2>&1; tail -3 "$out/ncu_attn_conv.log"
timeout -k 5 100 python tools/measure_dit.py --bf16 > "$out/measure_dit.log" 2>&1; tail -3 "$out/measure_dit.log"
ls -la "$out"
