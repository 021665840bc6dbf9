mkdir -p gpurun_out
for lz in 0 17 20 25 29 34 40; do
  if [ $lz -eq 0 ]; then unset FDFD_LZ; else export FDFD_LZ=$lz; fi
  python bench.py --steps 100 --warmup 5 --no-cpu --krylov-iters 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 lz', $lz, 'GDOF/s', round(d['value'],2))"
done
unset FDFD_LZ
python - <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, workloads
w = workloads.c3_phc_slab()
for lz in (0, 16, 22, 26, 32, 40):
    if lz: os.environ["FDFD_LZ"] = str(lz)
    A = workloads.make_operator(w, device=0)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(A.n, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
    y = torch.empty_like(x)
    A.bench_apply(x, y, warmup=5, iters=1)
    ms, _ = A.bench_apply(x, y, warmup=0, iters=50)
    print("C3 lz", lz, "GDOF/s", round(A.n / (ms / 50) / 1e6, 2))
    A.close()
PY
