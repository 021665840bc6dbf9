# round-2 GPU call 1 (1 GPU): time the opt-in paths written at the end of round 1, per-config numbers, material pipeline timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02c1_smi.txt
python bench.py > gpurun_out/r02c1_bench_line.json 2> gpurun_out/r02c1_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02c1_bench_line.json
FDFD_CORR_SKIP_ZERO=1 timeout 600 python -m pytest tests -m gpu -x -q -k "not matparams and not objects" > gpurun_out/r02c1_gpu_tests_skipz.log 2>&1; echo "gpu tests (skip-zero) rc=$?"; tail -2 gpurun_out/r02c1_gpu_tests_skipz.log
timeout 600 python scripts/bench_configs.py > gpurun_out/r02c1_configs_default.jsonl 2>&1; echo "configs rc=$?"
FDFD_CORR_SKIP_ZERO=1 timeout 600 python scripts/bench_configs.py > gpurun_out/r02c1_configs_skipz.jsonl 2>&1; echo "configs (skip-zero) rc=$?"
timeout 600 python scripts/bench_configs.py --skip-small --objects > gpurun_out/r02c1_configs_objects.jsonl 2>&1; echo "configs (objects) rc=$?"
FDFD_CORR_SKIP_ZERO=1 timeout 600 python scripts/bench_configs.py --skip-small --objects > gpurun_out/r02c1_configs_objects_skipz.jsonl 2>&1; echo "configs (objects, skip-zero) rc=$?"
for f in gpurun_out/r02c1_configs_*.jsonl; do echo $f; python - $f <<'PY'
import sys, json
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    if 'gdof_s' in d: print(' ', d['config'][:40], round(d['gdof_s'],1), 'frac', round(d['hbm_frac'],3), 'bpd', round(d['bytes_per_dof'],1), 'bicg', round(d['bicgstab_it_s'],1), 'setup', round(d['setup_s'],1))
PY
done
timeout 600 python scripts/bench_matparams.py > gpurun_out/r02c1_matparams_bench.jsonl 2>&1; echo "matparams bench rc=$?"; cut -c1-300 gpurun_out/r02c1_matparams_bench.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:matparams --csv --log-file gpurun_out/r02c1_matparams_launches.csv python scripts/bench_matparams.py > /dev/null 2>&1; echo "ncu rc=$?"; tail -8 gpurun_out/r02c1_matparams_launches.csv
