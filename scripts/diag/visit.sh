bash scripts/gpu_run.sh "tests:cli or host or graph or parity_gpu"
python bench.py --config c1 --steps 20 --warmup 3 --no-cpu-baseline --no-configs --recon-iters 50 > gpurun_out/c1.json 2> gpurun_out/c1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/c1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["recon"], d["check"])
PY
python scripts/diag/fn_eval_timing.py 2>&1 | tail -7
