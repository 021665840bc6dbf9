# round 2, call 18 (2 GPUs): halo data planes side by side - NCCL send/recv (default), copy-engine peer exchange
# (FDFD_PEER_HALO), in-place peer reads inside BiCGSTAB (FDFD_PEER_DIRECT), and the exchange overlapped with the apply
# (FDFD_HALO_OVERLAP) on top of NCCL capped at 2 CTAs / on top of the SM-free peer exchange; correctness first, then timing
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=r02c18
chk() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py > gpurun_out/${T}_dist_check_${N}_$tag.log 2>&1; echo "dist_check[$tag] rc=$?"; grep -E "DIST_CHECK" gpurun_out/${T}_dist_check_${N}_$tag.log | cut -c1-300; }
b() { tag=$1; shift; ( time env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 50 --warmup 5 $EXTRA ) > gpurun_out/${T}_scale_${N}_$tag.json 2> gpurun_out/${T}_scale_${N}_$tag.err; echo "bench[$tag] rc=$?"
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
    print('$tag: no line', e); import subprocess; print(subprocess.run("grep -m3 -E 'FdfdError|Error' gpurun_out/${T}_scale_${N}_$tag.err", shell=True, capture_output=True, text=True).stdout[:600])
PY
}
EXTRA="--no-configs --no-scale --no-single-call --no-cpu"
chk peer FDFD_PEER_HALO=1
b peer FDFD_PEER_HALO=1
chk direct FDFD_PEER_DIRECT=1
b direct FDFD_PEER_DIRECT=1
chk ovl_peer FDFD_HALO_OVERLAP=1 FDFD_PEER_HALO=1 FDFD_HALO_SM_RESERVE=0
b ovl_peer FDFD_HALO_OVERLAP=1 FDFD_PEER_HALO=1 FDFD_HALO_SM_RESERVE=0
b ovl_nccl2 FDFD_HALO_OVERLAP=1 NCCL_MAX_CTAS=2 FDFD_HALO_SM_RESERVE=4
EXTRA="--no-configs --no-cpu"
b default FDFD_NOP=1
