# round 2, call 16 (2 GPUs): z-slab correctness (one process per GPU and ONE process over both), exchange overlapped with
# the apply (in-kernel halo wait on the row-pair kernel) A/B, single-call boundary timing
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=r02c16
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/${T}_dist_tests_$N.log 2>&1; echo "dist tests rc=$?"; tail -3 gpurun_out/${T}_dist_tests_$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py > gpurun_out/${T}_dist_check_$N.log 2>&1; echo "dist_check rc=$?"; grep -E "DIST_CHECK" gpurun_out/${T}_dist_check_$N.log | cut -c1-300
b() { tag=$1; shift; ( time env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 50 --warmup 5 $EXTRA ) > gpurun_out/${T}_scale_${N}_$tag.json 2> gpurun_out/${T}_scale_${N}_$tag.err; echo "bench[$tag] rc=$?"; tail -2 gpurun_out/${T}_scale_${N}_$tag.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${T}_scale_${N}_$tag.json').read().strip().splitlines()[-1])
    print('$tag', 'N', d['n_gpus'], 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value'], 2))
    print('  parity', d['parity'].get('apply_rel_err'), d['parity'].get('traj_rel_err'), d['parity'].get('error'))
    print('  krylov', d['krylov']['iter_per_s'], d['krylov']['qmr_iter_per_s'], d['krylov']['error'])
    print('  halo', {k: d['halo'][k] for k in ('us', 'share_of_apply')} if d.get('halo') and 'us' in d['halo'] else d.get('halo'))
    print('  single_call', d.get('e2e_single_call'))
    for k in ('scale_c4', 'scale_c5'):
        c = d.get(k)
        if c: print(' ', k, {q: c.get(q) for q in ('gdof_s', 'hbm_frac', 'bicgstab_it_s', 'error')})
except Exception as e:
    print('$tag: no line', e)
PY
}
EXTRA=""
b overlap FDFD_HALO_OVERLAP=1
EXTRA="--no-configs --no-scale --no-single-call --no-cpu"
b serial FDFD_HALO_OVERLAP=0
b reserve2 FDFD_HALO_SM_RESERVE=2
b reserve8 FDFD_HALO_SM_RESERVE=8
