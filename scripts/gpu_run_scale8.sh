mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist$N rc=$?"; grep -E "DIST_CHECK" gpurun_out/dist_check_$N.log | head -3 | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 3 --master-addr 127.0.0.1 --master-port 29613 scripts/dist_check.py > gpurun_out/dist_check_3.log 2>&1; echo "dist3 rc=$?"; grep -E "DIST_CHECK" gpurun_out/dist_check_3.log | cut -c1-300
for n in 1 2 4 8; do
 if [ $n -gt $N ]; then continue; fi
 if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err;
 else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
 tail -1 gpurun_out/scale_$n.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'GDOF/s', round(d['value'],2), 'ms', round(d['ms_per_step'],4), 'it/s', round(d['krylov']['iter_per_s'],1), 'e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],3))" || tail -5 gpurun_out/scale_$n.err
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 100 --warmup 5 --dense-off > gpurun_out/scale_${N}_dense.json 2> gpurun_out/scale_${N}_dense.err
tail -1 gpurun_out/scale_${N}_dense.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('dense N', d['n_gpus'], 'GDOF/s', round(d['value'],2), 'ms', round(d['ms_per_step'],4), 'it/s', round(d['krylov']['iter_per_s'],1))"
