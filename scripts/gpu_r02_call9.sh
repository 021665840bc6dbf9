mkdir -p gpurun_out
run() { env "$@" timeout 300 python scripts/bench_k1.py $CFGS $CHK >> gpurun_out/r02c9_k1.jsonl 2>> gpurun_out/r02c9_k1.err; echo "[$*] rc=$?"; }
CFGS="c2 c3 c4"; CHK=""
run FDFD_RP_TMAP=1
CFGS="c2"; CHK="--no-check"
run FDFD_RP_NWC=5
for d in 1 2 4 7; do run FDFD_RP_DEBUG=$d; done
for d in 8 15; do run FDFD_B200_LIB=maxwellfdm.jl_b200/libfdfd_b200_abl.so FDFD_RP_DEBUG=$d; done
for c in 6 7 9 10; do run FDFD_RP_NCHUNK=$c; done
python - <<'PY'
import json
for l in open('gpurun_out/r02c9_k1.jsonl'):
    d = json.loads(l); print(d['tag'][:80].ljust(80), d['config'].ljust(8), d['ms'], d['gdof_s'], d['hbm_frac'], 'err', '%.1e' % d.get('rel_vs_general_kernel', -1), d.get('bicgstab_it_s'))
PY
tail -5 gpurun_out/r02c9_k1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowpair -s 3 -c 1 -o gpurun_out/r02c9_prof_rp python scripts/bench_k1.py c2 --no-check > gpurun_out/r02c9_ncu.log 2>&1; echo "ncu rc=$?"
