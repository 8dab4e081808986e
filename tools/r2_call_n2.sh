#!/bin/bash
mkdir -p gpurun_out/r2_n2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_n2/bench_n2.json 2> gpurun_out/r2_n2/bench_n2.err
tail -c 1500 gpurun_out/r2_n2/bench_n2.json; tail -3 gpurun_out/r2_n2/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2_n2/ref_n2.json 2> gpurun_out/r2_n2/ref_n2.err
tail -c 800 gpurun_out/r2_n2/ref_n2.json; tail -3 gpurun_out/r2_n2/ref_n2.err
