mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "apply or transpose or layout or config or large or golden" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in "" "--diag"; do
  python bench.py --steps 100 --warmup 5 --no-cpu $v 2>&1 | tail -1 > gpurun_out/tmp.json; cat gpurun_out/tmp.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['bytes_per_dof'], 'GDOF/s', round(d['value'],2), 'frac', round(d['roofline']['frac'],3), 'it/s', round(d['krylov']['iter_per_s'],1), d['clocks'])"
done
for lz in 16 25 34 50; do
  FDFD_LZ=$lz python bench.py --steps 50 --warmup 5 --no-cpu --krylov-iters 2 2>&1 | tail -1 > gpurun_out/tmp.json; cat gpurun_out/tmp.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('full lz', $lz, 'GDOF/s', round(d['value'],2))"
done
ncu --set full --clock-control none --import-source on -k regex:apply_tiled -s 3 -c 1 -o gpurun_out/prof_full_r01d \
    python bench.py --steps 3 --warmup 3 --no-cpu --krylov-iters 1 > gpurun_out/ncu_full.log 2>&1
compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; tail -4 gpurun_out/sanitizer.log
