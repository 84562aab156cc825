bash scripts/gpu_run.sh "tests:weights or gridding or gridded or host or cli"
GVM_GRID_TIMING=1 python bench.py --config c5 --scale 0.25 --steps 5 --warmup 3 --no-cpu-baseline --no-configs --recon-iters 0 > gpurun_out/c5q_timing.json 2> gpurun_out/c5q_timing.err
grep "gvm timing" gpurun_out/c5q_timing.err | head -60
python - <<PY
import json
d=json.loads(open("gpurun_out/c5q_timing.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["preprocessing"], d["check"])
PY
