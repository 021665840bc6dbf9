mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
FDFD_TY=8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "apply_all_boundary or layout or transpose or config or golden" > gpurun_out/pytest_ty8.log 2>&1; echo "pytest ty8 rc=$?"; tail -1 gpurun_out/pytest_ty8.log
python scripts/bench_configs.py > gpurun_out/configs_r01.jsonl 2> gpurun_out/configs_r01.err; python - <<'PY'
import json
for l in open('gpurun_out/configs_r01.jsonl'):
    d=json.loads(l)
    if 'gdof_s' in d: print(d['config'][:40], 'GDOF/s', round(d['gdof_s'],2), 'frac', round(d['hbm_frac'],3), 'off', round(d['offdiag_block_fraction'],4), 'bicg', round(d['bicgstab_it_s'],1), 'qmr', round(d['qmr_it_s'],1))
PY
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck.log 2>&1; tail -1 gpurun_out/racecheck.log
FDFD_TY=8 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1; tail -1 gpurun_out/memcheck.log
