python scripts/diag/noise_like_gradient.py > gpurun_out/noise_like_gradient.txt 2>&1; tail -8 gpurun_out/noise_like_gradient.txt
bash scripts/gpu_run.sh "tests:umma or full_size or smoke or gradient_vs_fp64 or error_maps or parity_reference or host"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["check"], d["recon"]["seconds"])
PY
