# round 2, call 25 (2 GPUs): single-call handle, standalone process: slab threads taking turns vs contending (A/B); then the
# bench line with the other ranks parked on a CPU barrier during the single-call block
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=r02c25
for tag in turns noturns; do
  if [ $tag = noturns ]; then export FDFD_NO_TURNS=1; else unset FDFD_NO_TURNS; fi
  timeout 600 python - > gpurun_out/${T}_single_call_${N}_$tag.json 2> gpurun_out/${T}_single_call_${N}_$tag.err <<PY
import json, sys
sys.path.insert(0, '.')
import bench
n = $N
print(json.dumps(bench.single_call_block((200, 200, 200 * n), (200, 200, 200), n, 5, 200)))
PY
  echo "single_call[$tag] rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/${T}_single_call_${N}_$tag.json').read().strip().splitlines()[-1]); print('$tag', 'apply', d['apply_gdof_s'], 'solve it/s', d['solve_iter_per_s'])"
done
unset FDFD_NO_TURNS
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu --no-scale ) > gpurun_out/${T}_scale_$N.json 2> gpurun_out/${T}_scale_$N.err; echo "bench rc=$?"; grep real gpurun_out/${T}_scale_$N.err
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_scale_$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', round(d['value'], 2), 'e2e', d['e2e']['value'], 'e2e_solve', d['e2e_solve']['iter_per_s'])
sc = d.get('e2e_single_call') or {}
print('single_call', {k: sc.get(k) for k in ('apply_gdof_s', 'solve_iter_per_s', 'error')})
PY
