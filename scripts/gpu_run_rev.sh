# mirrored (REV) tiled kernel + scalar identity mass: whole GPU suite, sanitizers on REV cases, HH bench, headline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(mirrored_arrangement_on_tiled_kernel and N4) or deep_grid" > gpurun_out/memcheck_rev.log 2>&1; tail -3 gpurun_out/memcheck_rev.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mirrored_arrangement_on_tiled_kernel and N4" > gpurun_out/racecheck_rev.log 2>&1; tail -3 gpurun_out/racecheck_rev.log
timeout 400 python scripts/bench_hh.py > gpurun_out/bench_hh.jsonl 2> gpurun_out/bench_hh.err; cat gpurun_out/bench_hh.jsonl | cut -c1-300; tail -3 gpurun_out/bench_hh.err
python bench.py --steps 100 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_c2.json; cut -c1-200 gpurun_out/bench_c2.json
python bench.py --steps 100 --warmup 5 --no-cpu --diag 2>&1 | tail -1 > gpurun_out/bench_diag.json; cut -c1-200 gpurun_out/bench_diag.json
