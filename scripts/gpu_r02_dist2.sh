mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_dist_tests_$N.log 2>&1; echo "dist tests rc=$?"; tail -3 gpurun_out/r02_dist_tests_$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py > gpurun_out/r02_dist_check_$N.log 2>&1; echo "dist_check rc=$?"; grep -E "DIST_CHECK" gpurun_out/r02_dist_check_$N.log | cut -c1-300
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 50 --warmup 5 ) > gpurun_out/r02_scale_$N.json 2> gpurun_out/r02_scale_$N.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_scale_$N.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r02_scale_$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
print('parity', d['parity']); print('krylov', d['krylov']); print('halo', d.get('halo')); print('e2e_solve', d['e2e_solve'])
print('c4', d.get('scale_c4')); print('c5', d.get('scale_c5'))
PY
