mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist$N rc=$?"; grep -E "DIST_CHECK" gpurun_out/dist_check_$N.log | head -3 | cut -c1-400; tail -2 gpurun_out/dist_check_$N.log | cut -c1-200
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 100 --warmup 5 --krylov-iters 20 > gpurun_out/dbg.json 2> gpurun_out/dbg.err; tail -1 gpurun_out/dbg.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$TAG', 'GDOF/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'it/s', round(d['krylov']['iter_per_s'],1))" || tail -3 gpurun_out/dbg.err; }
TAG=default run
TAG=no_prefetch FDFD_NO_HALO_PREFETCH=1 run
