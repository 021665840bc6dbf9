mkdir -p gpurun_out
run() { python bench.py --steps 60 --warmup 5 --no-cpu --krylov-iters 3 "$@" 2>&1 | tail -1 > gpurun_out/tmp.json; python -c "
import sys,json; d=json.load(open('gpurun_out/tmp.json')); print('$TAG', d['config']['bytes_per_dof'], 'GDOF/s', round(d['value'],2), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],2))"; }
TAG=ty16 run
TAG=ty16 run --diag
export FDFD_TY=8
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "apply_all_boundary or layout or host_and_device" > gpurun_out/pytest_ty8.log 2>&1; echo "pytest ty8 rc=$?"; tail -2 gpurun_out/pytest_ty8.log
TAG=ty8 run
TAG=ty8 run --diag
for lz in 12 20 30; do TAG="ty8 lz$lz" FDFD_LZ=$lz run --diag; done
