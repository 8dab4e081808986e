#!/bin/bash
# full GPU suite + the bench line (round-end state)
mkdir -p gpurun_out/r2_full
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r2_full/tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_full/bench.json 2> gpurun_out/r2_full/bench.err
tail -c 600 gpurun_out/r2_full/bench.json
