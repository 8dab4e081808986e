#!/bin/bash
# Round 2, measurement call after the hand-written tcgen05 GEMM became the only tensor-core path.  Everything lands in gpurun_out/r2_c2/.
set -u
out=gpurun_out/r2_c2
mkdir -p "$out"
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" > "$out/torch.txt" 2>&1
timeout -k 5 600 python -m pytest tests -q -m gpu -p no:cacheprovider > "$out/tests.log" 2>&1; tail -5 "$out/tests.log"
timeout -k 5 200 python tools/measure_dit.py --bf16 > "$out/measure_dit.log" 2>&1; tail -4 "$out/measure_dit.log"
timeout -k 5 300 python tools/measure_kernels.py --conv-tc --resblock --attn-ring > "$out/measure_kernels.log" 2>&1; tail -50 "$out/measure_kernels.log"
timeout -k 5 400 python tools/measure_configs.py --only prefill,caption32 > "$out/measure_configs.log" 2>&1; tail -12 "$out/measure_configs.log"
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:umma_kernel -s 2 -c 4 -o "$out/umma" -f python tools/profile_umma.py > "$out/ncu_umma.log" 2>&1; tail -3 "$out/ncu_umma.log"
timeout -k 5 400 python bench.py --steps 2 --warmup 3 > "$out/bench.json" 2> "$out/bench.err"; tail -c 3000 "$out/bench.json"; tail -5 "$out/bench.err"
ls -la "$out"
