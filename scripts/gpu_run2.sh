mkdir -p gpurun_out
python scripts/debug_qmr.py > gpurun_out/debug_qmr.log 2>&1; tail -12 gpurun_out/debug_qmr.log
# launch list of the bench (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --krylov-iters 3 > gpurun_out/bench_under_ncu.log 2>&1
# full capture of the apply kernel, full-tensor and diagonal variants
ncu --set full --clock-control none --import-source on -k regex:apply_tiled -s 3 -c 2 -o gpurun_out/prof_full_r01 \
    python bench.py --steps 3 --warmup 3 --no-cpu --krylov-iters 1 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:apply_tiled -s 3 -c 2 -o gpurun_out/prof_diag_r01 \
    python bench.py --steps 3 --warmup 3 --no-cpu --krylov-iters 1 --diag > gpurun_out/ncu_diag.log 2>&1
ls -la gpurun_out
python bench.py --steps 100 --warmup 5 --no-cpu --diag 2>&1 | tail -1
