mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 100 --warmup 5 --no-cpu --diag 2>&1 | tail -1 > gpurun_out/bench_diag.json; cat gpurun_out/bench_diag.json | cut -c1-400
python bench.py --steps 100 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_full.json; cat gpurun_out/bench_full.json | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:apply_tiled -s 3 -c 1 -o gpurun_out/prof_diag_r01b \
    python bench.py --steps 3 --warmup 3 --no-cpu --krylov-iters 1 --diag > gpurun_out/ncu_diag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:apply_tiled -s 3 -c 1 -o gpurun_out/prof_full_r01b \
    python bench.py --steps 3 --warmup 3 --no-cpu --krylov-iters 1 > gpurun_out/ncu_full.log 2>&1
