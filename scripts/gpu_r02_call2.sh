# round-2 GPU call 2: parity of the row-pair kernel on hardware, then A/B against the first-generation kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02c2_gpu_tests.log
run() { env "$@" timeout 300 python scripts/bench_k1.py c2 c3 c4 c5 >> gpurun_out/r02c2_k1.jsonl 2>> gpurun_out/r02c2_k1.err; echo "[$*] rc=$?"; }
run FDFD_K1_GEN=1
run FDFD_RP_NWC=11
run FDFD_RP_NWC=7 FDFD_RP_NST=6
run FDFD_RP_NWC=7 FDFD_RP_NST=4
python - <<'PY'
import json
for l in open('gpurun_out/r02c2_k1.jsonl'):
    d = json.loads(l); print(d['tag'][:40].ljust(40), d['config'].ljust(8), d['gdof_s'], d['hbm_frac'], 'err', '%.1e' % d.get('rel_vs_general_kernel', -1))
PY
tail -5 gpurun_out/r02c2_k1.err
