mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
FDFD_TY=16 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "apply_all_boundary or layout" > gpurun_out/pytest_ty16.log 2>&1; echo "pytest ty16 rc=$?"; tail -1 gpurun_out/pytest_ty16.log
FDFD_TY=8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "apply_all_boundary or layout" > gpurun_out/pytest_ty8.log 2>&1; echo "pytest ty8 rc=$?"; tail -1 gpurun_out/pytest_ty8.log
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck.log 2>&1; tail -2 gpurun_out/racecheck.log; grep -c "Race reported" gpurun_out/racecheck.log
FDFD_TY=8 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck8.log 2>&1; tail -1 gpurun_out/racecheck8.log
for v in "" "--diag"; do python bench.py --steps 100 --warmup 5 --no-cpu --krylov-iters 5 $v 2>&1 | tail -1 > gpurun_out/tmp.json; python -c "
import sys,json; d=json.load(open('gpurun_out/tmp.json')); print(d['config']['bytes_per_dof'], 'GDOF/s', round(d['value'],2), 'frac', round(d['roofline']['frac'],3))"; done
