mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c10_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02c10_gpu_tests.log
run() { env "$@" timeout 300 python scripts/bench_k1.py $CFGS $CHK >> gpurun_out/r02c10_k1.jsonl 2>> gpurun_out/r02c10_k1.err; echo "[$*] rc=$?"; }
CFGS="c2 c3 c4 c5"; CHK="--krylov"
run FDFD_RP_MDR=1
CFGS="c2 c4"; CHK="--no-check"
run FDFD_RP_MDR=0
for c in 6 7 8 9; do run FDFD_RP_NCHUNK=$c; done
for d in 7; do run FDFD_RP_DEBUG=$d; done
python - <<'PY'
import json
for l in open('gpurun_out/r02c10_k1.jsonl'):
    d = json.loads(l); print(d['tag'][:60].ljust(60), d['config'].ljust(8), d['ms'], d['gdof_s'], d['hbm_frac'], 'err', '%.1e' % d.get('rel_vs_general_kernel', -1), d.get('bicgstab_it_s'))
PY
tail -5 gpurun_out/r02c10_k1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowpair -s 3 -c 1 -o gpurun_out/r02c10_prof_rp_mdr python scripts/bench_k1.py c2 --no-check > gpurun_out/r02c10_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 12 --csv --log-file gpurun_out/r02c10_c3_launches.csv python scripts/bench_k1.py c3 --no-check > /dev/null 2>&1; tail -12 gpurun_out/r02c10_c3_launches.csv | cut -c1-250
