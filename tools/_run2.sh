set -x
timeout 900 python -m pytest tests/test_llm_gpu.py -x -q 2>&1 | tail -4
timeout 1500 python tools/measure_configs.py --only caption32 > gpurun_out/configs_r1c.jsonl 2> gpurun_out/configs_r1c.err; tail -3 gpurun_out/configs_r1c.err; cat gpurun_out/configs_r1c.jsonl | cut -c1-400
