# round 2, call 13: fused full-tensor shape of the row-pair kernel - parity on hardware, A/B against the two-pass plan
mkdir -p gpurun_out
T=r02c13
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or real_mass" > gpurun_out/${T}_gpu_tests_fused.log 2>&1; echo "fused tests rc=$?"; tail -3 gpurun_out/${T}_gpu_tests_fused.log
run() { env "$@" timeout 400 python scripts/bench_k1.py $CFGS $CHK >> gpurun_out/${T}_k1.jsonl 2>> gpurun_out/${T}_k1.err; echo "[$*] rc=$?"; }
CFGS="c2 c2d c3 c4 c5"; CHK="--krylov"
run FDFD_RP_FUSE_MIN=0.03
CHK="--no-check"
run FDFD_RP_FUSED=0
CFGS="c2 c3 c4"
run FDFD_RP_FUSE_MIN=0
python - <<'PY'
import json
for l in open('gpurun_out/r02c13_k1.jsonl'):
    d = json.loads(l); print(d['tag'][:40].ljust(40), d['config'].ljust(8), d['ms'], d['gdof_s'], d['hbm_frac'], d['bpd'], 'err', '%.1e' % d.get('rel_vs_general_kernel', -1), d.get('bicgstab_it_s'))
PY
tail -5 gpurun_out/${T}_k1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowpair -s 3 -c 1 -o gpurun_out/${T}_prof_rp_fused_c2d python scripts/bench_k1.py c2d --no-check > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowpair -s 3 -c 1 -o gpurun_out/${T}_prof_rp_fused_c3 python scripts/bench_k1.py c3 --no-check > gpurun_out/${T}_ncu3.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
