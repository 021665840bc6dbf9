# round 2, final 1-GPU pass with the final binary: smoke, the whole -m gpu suite, both bench arms, the ncu launch list of the
# bench command and one --set full capture of the dominant kernel
mkdir -p gpurun_out
T=r02f
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
( time python bench.py ) > gpurun_out/${T}_bench_line.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; grep real gpurun_out/${T}_bench.err
( time python bench.py --impl reference ) > gpurun_out/${T}_bench_ref_line.json 2> gpurun_out/${T}_bench_ref.err; echo "bench ref rc=$?"; grep real gpurun_out/${T}_bench_ref.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02f_bench_line.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('parity', d['parity'].get('apply_rel_err'), d['parity'].get('traj_rel_err')); print('krylov', d['krylov']['iter_per_s'], d['krylov']['qmr_iter_per_s'])
print('e2e_solve', d['e2e_solve']['iter_per_s']); print('single', (d.get('e2e_single_call') or {}).get('solve_iter_per_s'))
for c in d.get('configs', []): print({k: c.get(k) for k in ('config', 'gdof_s', 'hbm_frac', 'bicgstab_it_s', 'error')})
for k in ('scale_c4', 'scale_c5'): print(k, {q: d[k].get(q) for q in ('gdof_s', 'hbm_frac', 'bicgstab_it_s', 'error')})
r = json.loads(open('gpurun_out/r02f_bench_ref_line.json').read().strip().splitlines()[-1])
print('reference', r['value'], r['cpu_baseline']['cores'], r['krylov']['iter_per_s'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-configs --no-scale --no-single-call --no-cpu --krylov-iters 4 > gpurun_out/${T}_bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowpair -s 3 -c 1 -o gpurun_out/${T}_prof_rp_c2 python scripts/bench_k1.py c2 --no-check > gpurun_out/${T}_ncu.log 2>&1; echo "ncu full rc=$?"
