#!/bin/bash
# First GPU call of round 2: run what round 1 wrote but could not run any more, then measure it.
#   gpurun --timeout 1700 -- 'bash tools/r2_first_call.sh'   (the per-step limits add up to 1500 s; a clean run takes a few minutes)
# Everything lands in gpurun_out/r2_first/.
set -u
out=gpurun_out/r2_first
mkdir -p "$out"
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" > "$out/torch.txt" 2>&1   # pages torch in (about a minute on a fresh box)
UA2_RUN_UNVERIFIED=1 timeout -k 5 300 python -m pytest tests/test_zzz_unverified_gpu.py -q -m gpu -p no:cacheprovider > "$out/unverified_tests.log" 2>&1
tail -15 "$out/unverified_tests.log"
UA2_RUN_UNVERIFIED=1 timeout -k 5 400 compute-sanitizer --tool memcheck python -m pytest tests/test_zzz_unverified_gpu.py -q -m gpu -p no:cacheprovider \
    -k "conv_tc or resblock or attn_ring" > "$out/unverified_memcheck.log" 2>&1; tail -5 "$out/unverified_memcheck.log"
timeout -k 5 200 python tools/measure_dit.py --bf16 > "$out/measure_dit_bf16.log" 2>&1; tail -4 "$out/measure_dit_bf16.log"
timeout -k 5 200 python tools/measure_kernels.py --conv-tc --resblock --attn-ring > "$out/measure_kernels_options.log" 2>&1; tail -30 "$out/measure_kernels_options.log"
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:"ring_attn_kernel|sample_token_kernel|dit_attn_kernel" -s 6 -c 6 \
    -o "$out/new_kernels" -f python tools/profile_new_kernels.py > "$out/ncu_new_kernels.log" 2>&1; tail -3 "$out/ncu_new_kernels.log"
ls -la "$out"
