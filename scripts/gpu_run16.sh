mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
FDFD_TY=8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "solve or model_api" > gpurun_out/pytest_ty8.log 2>&1; echo "pytest ty8 rc=$?"; tail -1 gpurun_out/pytest_ty8.log
python scripts/debug_qmr.py 2>&1 | tail -8 | cut -c1-220
python scripts/bench_configs.py > gpurun_out/configs_r01.jsonl 2> gpurun_out/configs_r01.err; python - <<'PY'
import json
for l in open('gpurun_out/configs_r01.jsonl'):
    d=json.loads(l)
    if 'gdof_s' in d: print(d['config'][:40], 'GDOF/s', round(d['gdof_s'],2), 'frac', round(d['hbm_frac'],3), 'bicg', round(d['bicgstab_it_s'],1), 'qmr', round(d['qmr_it_s'],1))
    else: print({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in ('iters','true_relres','converged')}) for k,v in d.items()})
PY
FDFD_NO_DOT_FUSION=1 python bench.py --steps 20 --warmup 3 --no-cpu --krylov-iters 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('unfused it/s', round(d['krylov']['iter_per_s'],1))"
python bench.py --steps 20 --warmup 3 --no-cpu --krylov-iters 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fused it/s', round(d['krylov']['iter_per_s'],1))"
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck.log 2>&1; tail -1 gpurun_out/racecheck.log
