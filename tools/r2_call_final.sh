#!/bin/bash
# Last GPU call of the round: the whole GPU suite (with the tokenize composition tests) and bench.py's tokenize section on its own
mkdir -p gpurun_out/r2_final
timeout 200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_final/gpu_suite.log 2>&1
tail -30 gpurun_out/r2_final/gpu_suite.log
timeout 100 python -c "
import json, torch, bench
torch.cuda.set_device(0)
with torch.inference_mode():
    print(json.dumps(bench.bench_tokenize_frontends(torch.device('cuda', 0))))
" > gpurun_out/r2_final/bench_frontends.json 2> gpurun_out/r2_final/bench_frontends.err
tail -c 1500 gpurun_out/r2_final/bench_frontends.json; tail -3 gpurun_out/r2_final/bench_frontends.err
