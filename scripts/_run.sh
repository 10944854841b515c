for cs in 1 2; do
echo "=== cluster $cs"
CPB_TC_CLUSTER=$cs timeout 200 python scripts/diag_tc.py 2>&1 | grep "K=4096 pos\|K=576\|K=32 "
CPB_TC_CLUSTER=$cs timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bc$cs.json 2>gpurun_out/bc$cs.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bc$cs.json").read().strip().splitlines()[-1])
    print("CS $cs", round(d["value"]), d["ms_per_step"])
    g=d["roofline"]["groups_ms_per_step"]; print({k:round(v,2) for k,v in g.items() if v>0.8})
except Exception as e:
    print("CS $cs failed", e); print(open("gpurun_out/bc$cs.err").read()[-1500:])
PY
done
CPB_TC_CLUSTER=2 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
