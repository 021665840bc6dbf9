# quick check after a kernel edit: REV/HH parity tests, HH bench, headline bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "mirrored or hh_formulation or boundft or golden" 2>&1 | tail -2
timeout 400 python scripts/bench_hh.py > gpurun_out/bench_hh.jsonl 2> gpurun_out/bench_hh.err; cat gpurun_out/bench_hh.jsonl | cut -c1-220; tail -3 gpurun_out/bench_hh.err
python bench.py --steps 100 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_c2.json; cut -c1-200 gpurun_out/bench_c2.json
python bench.py --steps 100 --warmup 5 --no-cpu --diag 2>&1 | tail -1 > gpurun_out/bench_diag.json; cut -c1-200 gpurun_out/bench_diag.json
