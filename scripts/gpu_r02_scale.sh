# round 2: z-slab validation at N = all GPUs of the box with the default data plane (copy-engine peer exchange + exchange
# overlapped with the apply): correctness (one process per GPU, ONE process over all GPUs), then the full bench line
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=r02s
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py > gpurun_out/${T}_dist_check_$N.log 2>&1; echo "dist_check rc=$?"; grep -E "DIST_CHECK" gpurun_out/${T}_dist_check_$N.log | cut -c1-300
timeout 900 python scripts/multi_check.py $N > gpurun_out/${T}_multi_check_$N.log 2>&1; echo "multi_check rc=$?"; tail -2 gpurun_out/${T}_multi_check_$N.log | cut -c1-400
b() { tag=$1; shift; ( time env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 50 --warmup 5 $EXTRA ) > gpurun_out/${T}_scale_${N}_$tag.json 2> gpurun_out/${T}_scale_${N}_$tag.err; echo "bench[$tag] rc=$?"; grep real gpurun_out/${T}_scale_${N}_$tag.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${T}_scale_${N}_$tag.json').read().strip().splitlines()[-1])
    print('$tag', 'N', d['n_gpus'], 'value', round(d['value'], 2), 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value'], 2))
    print('  parity', d['parity'].get('apply_rel_err'), d['parity'].get('traj_rel_err'), d['parity'].get('error'))
    print('  krylov', d['krylov']['iter_per_s'], d['krylov']['qmr_iter_per_s'], d['krylov']['error'])
    print('  halo', {k: d['halo'][k] for k in ('us', 'share_of_apply', 'nvlink_frac')} if d.get('halo') and 'us' in d['halo'] else d.get('halo'))
    print('  e2e_solve', d.get('e2e_solve', {}).get('iter_per_s'))
    sc = d.get('e2e_single_call') or {}
    print('  single_call', {k: sc.get(k) for k in ('apply_gdof_s', 'apply_gdof_s_per_gpu', 'solve_iter_per_s', 'setup_s', 'error')})
    for k in ('scale_c4', 'scale_c5'):
        c = d.get(k)
        if c: print(' ', k, {q: c.get(q) for q in ('gdof_s', 'hbm_frac', 'bicgstab_it_s', 'error')})
except Exception as e:
    print('$tag: no line', e); import subprocess; print(subprocess.run("grep -m3 -E 'FdfdError|Error' gpurun_out/${T}_scale_${N}_$tag.err", shell=True, capture_output=True, text=True).stdout[:600])
PY
}
EXTRA="--no-cpu"
b default FDFD_NOP=1
EXTRA="--no-configs --no-scale --no-single-call --no-cpu"
b nccl FDFD_PEER_HALO=0
