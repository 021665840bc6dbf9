# round 2, call 30 (2 GPUs): last binary (longer bound on the in-kernel halo wait) - z-slab correctness and a short bench line
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=r02c30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py > gpurun_out/${T}_dist_check_$N.log 2>&1; echo "dist_check rc=$?"; grep -E "DIST_CHECK" gpurun_out/${T}_dist_check_$N.log | cut -c1-400
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu --no-configs --no-scale --no-single-call > gpurun_out/${T}_scale_$N.json 2> gpurun_out/${T}_scale_$N.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/${T}_scale_$N.json').read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'value', round(d['value'],2), 'parity', d['parity'].get('apply_rel_err'), d['parity'].get('traj_rel_err'), 'krylov', d['krylov']['iter_per_s'], 'plane', d['halo'].get('data_plane'))"
