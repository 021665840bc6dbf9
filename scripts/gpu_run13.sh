mkdir -p gpurun_out
python scripts/bench_configs.py > gpurun_out/configs_r01.jsonl 2> gpurun_out/configs_r01.err; cat gpurun_out/configs_r01.jsonl | cut -c1-700; tail -3 gpurun_out/configs_r01.err
