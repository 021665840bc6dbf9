# round 2, call 17 (1 GPU): occupancy flags fetched once per item (fused shape), single-call handle with ngpu = 1, full suite, bench
mkdir -p gpurun_out
T=r02c17
run() { env "$@" timeout 400 python scripts/bench_k1.py $CFGS $CHK >> gpurun_out/${T}_k1.jsonl 2>> gpurun_out/${T}_k1.err; echo "[$*] rc=$?"; }
CFGS="c2 c2d c3 c4 c5"; CHK="--krylov"
run FDFD_RP_DEBUG=0
CFGS="c2 c3 c4"; CHK="--no-check"
run FDFD_RP_FUSE_MIN=0
python - <<'PY'
import json
for l in open('gpurun_out/r02c17_k1.jsonl'):
    d = json.loads(l); print(d['tag'][:40].ljust(40), d['config'].ljust(8), d['ms'], d['gdof_s'], d['hbm_frac'], d['bpd'], 'err', '%.1e' % d.get('rel_vs_general_kernel', -1), d.get('bicgstab_it_s'))
PY
tail -5 gpurun_out/${T}_k1.err
timeout 600 python scripts/multi_check.py 1 > gpurun_out/${T}_multi_check.log 2>&1; echo "multi_check rc=$?"; tail -3 gpurun_out/${T}_multi_check.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
( time python bench.py ) > gpurun_out/${T}_bench_line.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02c17_bench_line.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'bpd', d['roofline']['bytes_per_dof'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
print('parity', d['parity']); print('krylov', d['krylov']); print('e2e_solve', d['e2e_solve']); print('single', d.get('e2e_single_call'))
for c in d.get('configs', []): print({k: c.get(k) for k in ('config', 'gdof_s', 'hbm_frac', 'bytes_per_dof', 'bicgstab_it_s', 'error')})
for k in ('scale_c4', 'scale_c5'): print(k, {q: d[k].get(q) for q in ('gdof_s', 'hbm_frac', 'bytes_per_dof', 'bicgstab_it_s', 'error')})
print('cpu', d.get('cpu_baseline'))
PY
( time python bench.py --impl reference ) > gpurun_out/${T}_bench_ref_line.json 2> gpurun_out/${T}_bench_ref.err; echo "bench ref rc=$?"; cut -c1-600 gpurun_out/${T}_bench_ref_line.json
