# round 2, call 14 (1 GPU): early stage release + table fill ahead of the item - parity on hardware, A/B timing
mkdir -p gpurun_out
T=r02c14
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or real_mass or boundary_conditions or boundft" > gpurun_out/${T}_gpu_tests_k1.log 2>&1; echo "k1 tests rc=$?"; tail -3 gpurun_out/${T}_gpu_tests_k1.log
run() { env "$@" timeout 400 python scripts/bench_k1.py $CFGS $CHK >> gpurun_out/${T}_k1.jsonl 2>> gpurun_out/${T}_k1.err; echo "[$*] rc=$?"; }
CFGS="c2 c2d c3 c4 c5"; CHK="--krylov"
run FDFD_RP_DEBUG=0
CHK="--no-check"
run FDFD_RP_DEBUG=16
run FDFD_RP_DEBUG=32
CFGS="c2 c3 c4"
run FDFD_RP_FUSE_MIN=0
run FDFD_RP_FUSED=0
python - <<'PY'
import json
for l in open('gpurun_out/r02c14_k1.jsonl'):
    d = json.loads(l); print(d['tag'][:40].ljust(40), d['config'].ljust(8), d['ms'], d['gdof_s'], d['hbm_frac'], d['bpd'], 'err', '%.1e' % d.get('rel_vs_general_kernel', -1), d.get('bicgstab_it_s'))
PY
tail -5 gpurun_out/${T}_k1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowpair -s 3 -c 1 -o gpurun_out/${T}_prof_rp_c2 python scripts/bench_k1.py c2 --no-check > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowpair -s 3 -c 1 -o gpurun_out/${T}_prof_rp_c5 python scripts/bench_k1.py c5 --no-check > gpurun_out/${T}_ncu5.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
( time python bench.py ) > gpurun_out/${T}_bench_line.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
