mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck.log 2>&1; tail -1 gpurun_out/racecheck.log
timeout 900 python scripts/bench_configs.py --c4 > gpurun_out/configs_r01.jsonl 2> gpurun_out/configs_r01.err; tail -2 gpurun_out/configs_r01.err; python - <<'PY'
import json
for l in open('gpurun_out/configs_r01.jsonl'):
    d=json.loads(l)
    if 'gdof_s' in d: print(d['config'][:40], 'GDOF/s', round(d['gdof_s'],2), 'ms', round(d['ms_per_apply'],3), 'frac', round(d['hbm_frac'],3), 'off', round(d['offdiag_block_fraction'],4), 'bicg', round(d['bicgstab_it_s'],1), 'qmr', round(d['qmr_it_s'],1))
PY
