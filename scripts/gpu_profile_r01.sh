# Round-1 profile capture (1 GPU): launch list of the bench command + ncu --set full of the dominant kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu --krylov-iters 5 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"apply_tiled|offdiag_march" -s 6 -c 2 -o gpurun_out/prof_c2_r01 \
    python bench.py --steps 3 --warmup 3 --no-cpu --krylov-iters 1 > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:apply_tiled -s 3 -c 1 -o gpurun_out/prof_dense_r01 \
    python bench.py --steps 3 --warmup 3 --no-cpu --krylov-iters 1 --dense-off > gpurun_out/ncu_dense.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:apply_tiled -s 3 -c 1 -o gpurun_out/prof_hh_r01 \
    python scripts/bench_hh.py --variant 1 > gpurun_out/ncu_hh.log 2>&1
ncu --set full --clock-control none -k regex:"k_xr|k_p|k_s|k_dot" -s 10 -c 5 -o gpurun_out/prof_krylov_r01 \
    python bench.py --steps 3 --warmup 3 --no-cpu --krylov-iters 4 > gpurun_out/ncu_krylov.log 2>&1
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -1 gpurun_out/bench_r01.json | cut -c1-600
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_r01.json 2>&1; tail -1 gpurun_out/bench_ref_r01.json | cut -c1-300
