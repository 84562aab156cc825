SECONDS=0
python bench.py > gpurun_out/bench_default_.json 2> gpurun_out/bench_default_.err
echo "default bench wall seconds: $SECONDS"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_default_.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
for k,v in d["configs"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("roofline",{}).get("frac"), v.get("error"))
PY
