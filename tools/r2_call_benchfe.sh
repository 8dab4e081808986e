#!/bin/bash
# bench.py's tokenize front-end section on its own (the whole bench line is the driver's round-end run)
mkdir -p gpurun_out/r2_frontend4
timeout 240 python -c "
import json, torch, bench
torch.cuda.set_device(0)
with torch.inference_mode():
    print(json.dumps(bench.bench_tokenize_frontends(torch.device('cuda', 0))))
" > gpurun_out/r2_frontend4/bench_frontends.json 2> gpurun_out/r2_frontend4/bench_frontends.err
tail -c 3000 gpurun_out/r2_frontend4/bench_frontends.json; tail -5 gpurun_out/r2_frontend4/bench_frontends.err
