mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for v in "" "--diag" "--dense-off"; do python bench.py --steps 100 --warmup 5 --no-cpu --krylov-iters 5 $v 2>&1 | tail -1 > gpurun_out/tmp.json; python -c "
import sys,json; d=json.load(open('gpurun_out/tmp.json')); print('$v', round(d['config']['bytes_per_dof'],2), 'GDOF/s', round(d['value'],2), 'frac', round(d['roofline']['frac'],3), 'offfrac', round(d['config']['offdiag_block_fraction'],4), 'it/s', round(d['krylov']['iter_per_s'],1))"; done
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck.log 2>&1; tail -1 gpurun_out/racecheck.log
ncu --set full --clock-control none --import-source on -k regex:apply_tiled -s 3 -c 1 -o gpurun_out/prof_full_r01f \
    python bench.py --steps 3 --warmup 3 --no-cpu --krylov-iters 1 > gpurun_out/ncu_full.log 2>&1
