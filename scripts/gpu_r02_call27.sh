# round 2, call 27 (1 GPU): one tile row per compute warp (10 / 14 compute warps) against the row-pair form
mkdir -p gpurun_out
T=r02c27
run() { env "$@" timeout 400 python scripts/bench_k1.py $CFGS $CHK >> gpurun_out/${T}_k1.jsonl 2>> gpurun_out/${T}_k1.err; echo "[$*] rc=$?"; }
CFGS="c2 c4"; CHK="--krylov"
run FDFD_RP_RPW1=0
run FDFD_RP_RPW1=10
run FDFD_RP_RPW1=14
python - <<'PY'
import json
for l in open('gpurun_out/r02c27_k1.jsonl'):
    d = json.loads(l); print(d['tag'][:40].ljust(40), d['config'].ljust(8), d['ms'], d['gdof_s'], d['hbm_frac'], d['bpd'], 'err', '%.1e' % d.get('rel_vs_general_kernel', -1), d.get('bicgstab_it_s'))
PY
tail -5 gpurun_out/${T}_k1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowpair -s 3 -c 1 -o gpurun_out/${T}_prof_rpw10_c2 env FDFD_RP_RPW1=10 python scripts/bench_k1.py c2 --no-check > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
