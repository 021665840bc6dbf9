# round 2, call 20 (1 GPU): where does the base cost of the fused full-tensor shape come from?  A/B: the real-row shape on a
# 4-stage ring (FDFD_RP_NST4), the fused shape with its off-diagonal arithmetic compiled in but skipped (FDFD_RP_DEBUG=128)
mkdir -p gpurun_out
T=r02c20
run() { env "$@" timeout 400 python scripts/bench_k1.py $CFGS $CHK >> gpurun_out/${T}_k1.jsonl 2>> gpurun_out/${T}_k1.err; echo "[$*] rc=$?"; }
CFGS="c2 c4"; CHK="--no-check"
run FDFD_RP_DEBUG=0
run FDFD_RP_NST4=1
run FDFD_RP_FUSE_MIN=0
run FDFD_RP_FUSE_MIN=0 FDFD_RP_DEBUG=128
run FDFD_RP_FUSE_MIN=0 FDFD_RP_DEBUG=16
python - <<'PY'
import json
for l in open('gpurun_out/r02c20_k1.jsonl'):
    d = json.loads(l); print(d['tag'][:50].ljust(50), d['config'].ljust(8), d['ms'], d['gdof_s'], d['hbm_frac'], d['bpd'])
PY
tail -5 gpurun_out/${T}_k1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
