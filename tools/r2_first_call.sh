#!/bin/bash
# First GPU call of round 2: run what round 1 wrote but could not run any more, then measure it.
#   gpurun --timeout 2500 -- 'bash tools/r2_first_call.sh'   (the per-step limits add up to 2400 s; a clean run takes a few minutes)
# Everything lands in gpurun_out/r2_first/.
set -u
out=gpurun_out/r2_first
mkdir -p "$out"
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" > "$out/torch.txt" 2>&1   # pages torch in (about a minute on a fresh box)
UA2_RUN_UNVERIFIED=1 timeout -k 5 300 python -m pytest tests/test_zzz_unverified_gpu.py -q -m gpu -p no:cacheprovider > "$out/unverified_tests.log" 2>&1
tail -15 "$out/unverified_tests.log"
UA2_RUN_UNVERIFIED=1 timeout -k 5 240 compute-sanitizer --tool memcheck python -m pytest tests/test_zzz_unverified_gpu.py -q -m gpu -p no:cacheprovider \
    -k "conv_tc or resblock or attn_ring" > "$out/unverified_memcheck.log" 2>&1; tail -5 "$out/unverified_memcheck.log"
timeout -k 5 200 python tools/measure_dit.py --bf16 > "$out/measure_dit_bf16.log" 2>&1; tail -4 "$out/measure_dit_bf16.log"
timeout -k 5 200 python tools/measure_kernels.py --conv-tc --resblock --attn-ring > "$out/measure_kernels_options.log" 2>&1; tail -30 "$out/measure_kernels_options.log"
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:"ring_attn_kernel|sample_token_kernel|dit_attn_kernel" -s 6 -c 6 \
    -o "$out/new_kernels" -f python tools/profile_new_kernels.py > "$out/ncu_new_kernels.log" 2>&1; tail -3 "$out/ncu_new_kernels.log"
# A/B of the codec section of bench.py with the two codec options on (both runs in this call so that they share a box)
for opts in "" "conv_tc=1,resblock_fused=1"; do
  UA2_OPTIONS="$opts" timeout -k 5 240 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-flow-decoder > "$out/bench_codec_${opts:-default}.json" 2> "$out/bench_codec_${opts:-default}.err"
  python - "$out/bench_codec_${opts:-default}.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", d.get("value"), "codec", {k: d.get("codec", {}).get(k) for k in ("encode_ms", "decode_ms", "rtf_x_realtime")})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
ls -la "$out"
