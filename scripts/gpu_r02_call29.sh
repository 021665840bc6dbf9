# round 2, call 29 (2 GPUs): the final binary on z-slabs - dist_check (incl. deep slabs) and the bench line
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=r02c29
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py > gpurun_out/${T}_dist_check_$N.log 2>&1; echo "dist_check rc=$?"; grep -E "DIST_CHECK" gpurun_out/${T}_dist_check_$N.log | cut -c1-400
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu ) > gpurun_out/${T}_scale_$N.json 2> gpurun_out/${T}_scale_$N.err; echo "bench rc=$?"; grep real gpurun_out/${T}_scale_$N.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_scale_$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 3), 'e2e', d['e2e']['value'])
print('  parity', d['parity'].get('apply_rel_err'), d['parity'].get('traj_rel_err'), d['parity'].get('error'))
print('  krylov', d['krylov']['iter_per_s'], d['krylov']['qmr_iter_per_s'], d['krylov']['error'])
print('  halo', {k: d['halo'].get(k) for k in ('us', 'share_of_apply', 'data_plane')})
print('  e2e_solve', d.get('e2e_solve', {}).get('iter_per_s'))
sc = d.get('e2e_single_call') or {}
print('  single_call', {k: sc.get(k) for k in ('apply_gdof_s', 'solve_iter_per_s', 'error')})
for k in ('scale_c4', 'scale_c5'):
    c = d.get(k)
    if c: print(' ', k, {q: c.get(q) for q in ('gdof_s', 'hbm_frac', 'bytes_per_dof', 'bicgstab_it_s', 'error')})
PY
