CPB_TC_CLUSTER=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bc1.json 2>gpurun_out/bc1.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bc1.json").read().strip().splitlines()[-1])
    print(round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]))
    g=d["roofline"]["groups_ms_per_step"]; print({k:round(v,2) for k,v in g.items() if v>0.5})
except Exception as e:
    print("failed", e); print(open("gpurun_out/bc1.err").read()[-1500:])
PY
CPB_TC_CLUSTER=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
