mkdir -p gpurun_out
run() { python bench.py --steps 60 --warmup 5 --no-cpu --krylov-iters 3 "$@" 2>&1 | tail -1 > gpurun_out/tmp.json; python -c "
import sys,json; d=json.load(open('gpurun_out/tmp.json')); print('$TAG', d['config']['bytes_per_dof'], 'GDOF/s', round(d['value'],2), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],2))"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
FDFD_TY=16 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "apply_all_boundary or layout" > gpurun_out/pytest_ty16.log 2>&1; echo "pytest ty16 rc=$?"; tail -2 gpurun_out/pytest_ty16.log
FDFD_TY=8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "apply_all_boundary or layout" > gpurun_out/pytest_ty8.log 2>&1; echo "pytest ty8 rc=$?"; tail -2 gpurun_out/pytest_ty8.log
TAG=default run
TAG=default run --diag
TAG=ty16 FDFD_TY=16 run --diag
TAG=ty8 FDFD_TY=8 run
compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; tail -2 gpurun_out/sanitizer.log
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck.log 2>&1; tail -3 gpurun_out/racecheck.log
ncu --set full --clock-control none --import-source on -k regex:apply_tiled -s 3 -c 1 -o gpurun_out/prof_diag_r01e \
    python bench.py --steps 3 --warmup 3 --no-cpu --krylov-iters 1 --diag > gpurun_out/ncu_diag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:apply_tiled -s 3 -c 1 -o gpurun_out/prof_full_r01e \
    python bench.py --steps 3 --warmup 3 --no-cpu --krylov-iters 1 > gpurun_out/ncu_full.log 2>&1
