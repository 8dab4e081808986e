set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-codec > gpurun_out/b_new.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/b_new.log | head -1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-codec --attn-direct 1 > gpurun_out/b_new_ad1.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/b_new_ad1.log | head -1
timeout 900 python bench.py > gpurun_out/bench_r1_full2.log 2>&1; tail -1 gpurun_out/bench_r1_full2.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_v3b.csv python tools/profile_frame.py --frames 3 > gpurun_out/prof_stdout5.log 2>&1
tail -n 2 gpurun_out/prof_stdout5.log
