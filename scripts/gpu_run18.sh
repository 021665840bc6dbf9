mkdir -p gpurun_out
for lib in libfdfd_b200.so libfdfd_b200_u2.so libfdfd_b200_u4.so; do
  export FDFD_B200_LIB=$PWD/maxwellfdm.jl_b200/$lib
  for v in "" "--diag" "--dense-off"; do python bench.py --steps 100 --warmup 5 --no-cpu --krylov-iters 10 $v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$v', 'GDOF/s', round(d['value'],2), 'frac', round(d['roofline']['frac'],3), 'it/s', round(d['krylov']['iter_per_s'],1))"; done
done
export FDFD_B200_LIB=$PWD/maxwellfdm.jl_b200/libfdfd_b200_u2.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "apply_all_boundary or layout or transpose or config" > gpurun_out/pytest_u2.log 2>&1; echo "pytest u2 rc=$?"; tail -1 gpurun_out/pytest_u2.log
