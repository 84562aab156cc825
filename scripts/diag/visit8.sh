SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$? wall=$SECONDS s"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["check"], d["recon"]["seconds"])
for k,v in d["configs"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("e2e",{}).get("value"), v.get("roofline",{}).get("frac"), v.get("preprocessing"), v.get("error"))
PY
tail -3 gpurun_out/bench_n8.err
