# Round-2 first GPU calls (nothing below has run on hardware yet - it was written after the round-1 GPU budget was spent
# and checked only under the CPU logic-check build, tests/emu/).
#   1 GPU :  gpurun --timeout 1200 -- 'bash scripts/gpu_r02_first.sh single'
#   2 GPUs:  gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_r02_first.sh peer'
mkdir -p gpurun_out
case "$1" in
single)
  # new GPU tests first (material pipeline), then the whole suite, the bench line, the material-pipeline timing + ncu
  timeout 600 python -m pytest tests -m gpu -x -q -k "matparams or objects or reduced or trajectory or edge_cases" > gpurun_out/r02_matparams_tests.log 2>&1; echo "matparams tests rc=$?"; tail -3 gpurun_out/r02_matparams_tests.log
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02_gpu_tests.log
  python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_line.json
  # BiCGSTAB with sigma accumulated by the p update (default since the end of round 1) vs the separate (rhat, v) pass
  FDFD_BICGSTAB_CLASSIC=1 python bench.py --no-cpu > gpurun_out/r02_bench_line_classic_bicgstab.json 2>> gpurun_out/r02_bench.err; echo "bench (classic BiCGSTAB) rc=$?"
  for f in gpurun_out/r02_bench_line.json gpurun_out/r02_bench_line_classic_bicgstab.json; do python -c "
import sys,json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', 'BiCGSTAB it/s', round(d['krylov']['iter_per_s'],1), 'hbm_frac', round(d['krylov']['hbm_frac'],3))"; done
  # correction pass that skips exact zeros (opt-in): parity of the whole suite with it on, then C2-C5 with / without
  FDFD_CORR_SKIP_ZERO=1 timeout 900 python -m pytest tests -m gpu -x -q -k "not matparams and not objects" > gpurun_out/r02_gpu_tests_skipz.log 2>&1; echo "gpu tests (skip-zero) rc=$?"; tail -2 gpurun_out/r02_gpu_tests_skipz.log
  timeout 900 python scripts/bench_configs.py > gpurun_out/r02_configs_default.jsonl 2>&1; echo "configs rc=$?"
  FDFD_CORR_SKIP_ZERO=1 timeout 900 python scripts/bench_configs.py > gpurun_out/r02_configs_skipz.jsonl 2>&1; echo "configs (skip-zero) rc=$?"
  for f in gpurun_out/r02_configs_default.jsonl gpurun_out/r02_configs_skipz.jsonl; do echo $f; cut -c1-200 $f | grep gdof_s; done
  timeout 900 python scripts/bench_configs.py --skip-small --objects > gpurun_out/r02_configs_objects.jsonl 2>&1; echo "configs (objects) rc=$?"; cut -c1-260 gpurun_out/r02_configs_objects.jsonl
  FDFD_CORR_SKIP_ZERO=1 timeout 900 python scripts/bench_configs.py --skip-small --objects > gpurun_out/r02_configs_objects_skipz.jsonl 2>&1; echo "configs (objects, skip-zero) rc=$?"; cut -c1-260 gpurun_out/r02_configs_objects_skipz.jsonl
  timeout 600 python scripts/bench_matparams.py > gpurun_out/r02_matparams_bench.jsonl 2>&1; echo "matparams bench rc=$?"; cat gpurun_out/r02_matparams_bench.jsonl | cut -c1-300
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:matparams --csv --log-file gpurun_out/r02_matparams_launches.csv python scripts/bench_matparams.py > /dev/null 2>&1; echo "ncu rc=$?"
  ;;
peer)
  N=$(nvidia-smi -L | wc -l)
  # correctness of the SM-free peer-memory halo exchange (csrc/peer.cpp), then apply / BiCGSTAB throughput with and without it
  for mode in "" "FDFD_PEER_HALO=1" "FDFD_PEER_DIRECT=1"; do
    env $mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py > gpurun_out/r02_dist_check_${N}_${mode%%=*}.log 2>&1; echo "dist_check [$mode] rc=$?"; grep -E "DIST_CHECK" gpurun_out/r02_dist_check_${N}_${mode%%=*}.log | cut -c1-300
    env $mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r02_scale_${N}_${mode%%=*}.json 2> gpurun_out/r02_scale_${N}_${mode%%=*}.err; echo "bench [$mode] rc=$?"
    tail -1 gpurun_out/r02_scale_${N}_${mode%%=*}.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'GDOF/s', round(d['value'],2), 'ms', round(d['ms_per_step'],4), 'it/s', round(d['krylov']['iter_per_s'],1))" || tail -5 gpurun_out/r02_scale_${N}_${mode%%=*}.err
  done
  # the in-kernel halo wait on top of the SM-free exchange (the combination the round-1 experiment was missing)
  env FDFD_PEER_HALO=1 FDFD_INKERNEL_HALO_WAIT=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29623 scripts/dist_inkernel_check.py > gpurun_out/r02_inkernel_peer_$N.log 2>&1; echo "inkernel+peer rc=$?"; tail -3 gpurun_out/r02_inkernel_peer_$N.log | cut -c1-300
  ;;
*) echo "usage: $0 single|peer";;
esac
