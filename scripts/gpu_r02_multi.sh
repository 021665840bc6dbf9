# round 2: the single-call handle (ONE process, host threads = GPUs) on all GPUs of the box: correctness, then apply / solve rates
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T=r02m
timeout 600 python scripts/multi_check.py $N > gpurun_out/${T}_multi_check_$N.log 2>&1; echo "multi_check rc=$?"; tail -2 gpurun_out/${T}_multi_check_$N.log | cut -c1-300
timeout 600 python - > gpurun_out/${T}_single_call_$N.json 2> gpurun_out/${T}_single_call_$N.err <<PY
import json, sys
sys.path.insert(0, '.')
import bench
n = $N
print(json.dumps(bench.single_call_block((200, 200, 200 * n), (200, 200, 200), n, 5, 200)))
PY
echo "single_call rc=$?"; cut -c1-700 gpurun_out/${T}_single_call_$N.json; tail -2 gpurun_out/${T}_single_call_$N.err | cut -c1-300
