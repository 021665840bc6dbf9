# round-2 GPU call 3: where does the row-pair kernel's time go?  ablations (timing only) + one ncu --set full capture
mkdir -p gpurun_out
run() { env "$@" timeout 300 python scripts/bench_k1.py c2 --no-check >> gpurun_out/r02c3_k1.jsonl 2>> gpurun_out/r02c3_k1.err; echo "[$*] rc=$?"; }
for d in 0 1 2 4 3 5 6 7; do run FDFD_RP_NWC=7 FDFD_RP_NST=6 FDFD_RP_DEBUG=$d; done
for d in 0 1 7; do run FDFD_RP_NWC=11 FDFD_RP_DEBUG=$d; done
for g in 74 111 148; do run FDFD_RP_NWC=7 FDFD_RP_NST=6 FDFD_RP_GRID=$g; done
for c in 4 6 10 13; do run FDFD_RP_NWC=7 FDFD_RP_NST=6 FDFD_RP_NCHUNK=$c; done
python - <<'PY'
import json
for l in open('gpurun_out/r02c3_k1.jsonl'):
    d = json.loads(l); print(d['tag'][:60].ljust(60), d['config'].ljust(8), d['ms'], d['gdof_s'], d['hbm_frac'])
PY
FDFD_RP_NWC=7 FDFD_RP_NST=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowpair -s 3 -c 1 -o gpurun_out/r02c3_prof_rp7 python scripts/bench_k1.py c2 --no-check > gpurun_out/r02c3_ncu.log 2>&1; echo "ncu rc=$?"
FDFD_RP_NWC=11 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowpair -s 3 -c 1 -o gpurun_out/r02c3_prof_rp11 python scripts/bench_k1.py c2 --no-check >> gpurun_out/r02c3_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
