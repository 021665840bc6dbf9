# Round-1 final verification on 1 GPU: full -m gpu suite, sanitizers, bench lines, profiles.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck.log 2>&1; tail -1 gpurun_out/racecheck.log
compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1; tail -1 gpurun_out/memcheck.log
bash scripts/gpu_profile_r01.sh > gpurun_out/profile_run.log 2>&1; tail -2 gpurun_out/profile_run.log | cut -c1-400
timeout 900 python scripts/bench_configs.py --c4 --c5 > gpurun_out/configs_r01.jsonl 2> gpurun_out/configs_r01.err; python - <<'PY'
import json
for l in open('gpurun_out/configs_r01.jsonl'):
    d=json.loads(l)
    if 'gdof_s' in d: print(d['config'][:40], 'GDOF/s', round(d['gdof_s'],2), 'frac', round(d['hbm_frac'],3), 'bicg', round(d['bicgstab_it_s'],1), 'qmr', round(d['qmr_it_s'],1))
PY
for v in "--diag" "--dense-off"; do python bench.py --steps 100 --warmup 5 --no-cpu $v 2>&1 | tail -1 > gpurun_out/bench_r01$v.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r01$v.json')); print('$v', 'GDOF/s', round(d['value'],2), 'frac', round(d['roofline']['frac'],3), 'it/s', round(d['krylov']['iter_per_s'],1))"; done
timeout 400 python scripts/bench_hh.py > gpurun_out/bench_hh.jsonl 2> gpurun_out/bench_hh.err; cut -c1-160 gpurun_out/bench_hh.jsonl
