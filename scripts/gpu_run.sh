#!/bin/bash
# One GPU visit = a list of stages, e.g.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_run.sh tests golden bench refarm'
# Every stage writes its full log under gpurun_out/ (merged back by gpurun) and prints a short tail.
#   tests[:expr]   pytest -m gpu (optionally -k expr)
#   golden         regenerate the reference-built golden vectors (tests/golden/make_golden.py)
#   bench[:args]   bench.py with the given extra args ("," separates them), JSON line -> gpurun_out/bench*.json
#   refarm[:args]  bench.py --impl reference
#   refsweep       reference arm at 2e5 and 1e6 visibilities per step (the full 1e7 point is refarm's full_workload_check)
#   launches[:cfg] ncu launch list (gpu__time_duration) of a short bench run
#   ncu:<kernel regex>[:cfg[:skip]]   one ncu --set full capture of that kernel
#   multi:N        test_multi_gpu + torchrun bench at N ranks (use with gpurun --gpus N)
#   sanitizer      compute-sanitizer memcheck over the small parity tests
#   diag:<name>[:args]   build scripts/diag/<name>.cu for sm_100a on the box and run it (umma_arith, umma_fp8mix, mufu_err,
#                  cluster_occupancy) -> gpurun_out/diag_<name>.txt
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
for stage in "$@"; do
  name=${stage%%:*}; arg=""; [[ "$stage" == *:* ]] && arg=${stage#*:}
  echo "=== stage $stage"
  case $name in
    tests)
      timeout 1700 python -m pytest tests -m gpu -q -s ${arg:+-k "$arg"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
      grep -n "^\[\|passed\|failed\|Error\|rc=" gpurun_out/pytest_gpu.log | tail -n 40 ;;
    golden)
      for which in "" ext priors; do timeout 600 python tests/golden/make_golden.py $which > gpurun_out/golden_$which.log 2>&1; echo "golden '$which' rc=$?"; done ;;
    bench)
      tag=$(echo "${arg:-default}" | tr -c 'A-Za-z0-9' '_'); extra=$(echo "$arg" | tr ',' ' ')
      timeout 1500 python bench.py $extra > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
      tail -c 1500 gpurun_out/bench_$tag.json; tail -n 3 gpurun_out/bench_$tag.err ;;
    refarm)
      tag=$(echo "${arg:-default}" | tr -c 'A-Za-z0-9' '_'); extra=$(echo "$arg" | tr ',' ' ')
      timeout 1700 python bench.py --impl reference $extra > gpurun_out/refarm_$tag.json 2> gpurun_out/refarm_$tag.err; echo "refarm rc=$?"
      tail -c 1200 gpurun_out/refarm_$tag.json ;;
    refsweep)
      for z in 200000 1000000; do
        timeout 900 python bench.py --impl reference --steps 3 --ref-sample $z --no-full-check --no-ref-recon > gpurun_out/refsweep_$z.json 2> gpurun_out/refsweep_$z.err; echo "refsweep $z rc=$?"
        tail -c 400 gpurun_out/refsweep_$z.json
      done ;;
    launches)
      cfg=${arg:-c2}
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$cfg.csv \
        python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-configs --recon-iters 0 > gpurun_out/launches_$cfg.log 2>&1; echo "launches rc=$?"
      python scripts/launch_summary.py gpurun_out/launches_$cfg.csv > gpurun_out/launches_${cfg}_summary.txt 2>&1; head -n 30 gpurun_out/launches_${cfg}_summary.txt ;;
    ncu)
      IFS=: read -r kern cfg skip <<< "$arg"; cfg=${cfg:-c2}; skip=${skip:-2}
      timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$kern" -s $skip -c 1 -f -o gpurun_out/ncu_${kern}_$cfg \
        python bench.py --config $cfg --steps 1 --warmup 3 --no-cpu-baseline --no-configs --recon-iters 0 > gpurun_out/ncu_${kern}_$cfg.log 2>&1; echo "ncu rc=$?"
      tail -n 3 gpurun_out/ncu_${kern}_$cfg.log ;;
    multi)
      N=${arg:-2}
      timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -s > gpurun_out/pytest_multi_gpu.log 2>&1; echo "multi rc=$?" >> gpurun_out/pytest_multi_gpu.log
      grep -n "^\[\|passed\|failed\|skipped\|rc=" gpurun_out/pytest_multi_gpu.log | tail
      timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
        bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
      tail -c 1500 gpurun_out/bench_n$N.json ;;
    sanitizer)
      timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_edges_gpu.py -m gpu -q -x \
        -k "not full_size" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
      tail -n 5 gpurun_out/sanitizer_memcheck.log ;;
    diag)
      IFS=: read -r prog pargs <<< "$arg"
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/diag_$prog scripts/diag/$prog.cu > gpurun_out/diag_$prog.txt 2>&1 &&
        timeout 300 /tmp/diag_$prog $pargs >> gpurun_out/diag_$prog.txt 2>&1; echo "diag $prog rc=$?"
      tail -n 25 gpurun_out/diag_$prog.txt ;;
    *) echo "unknown stage $stage" ;;
  esac
done
