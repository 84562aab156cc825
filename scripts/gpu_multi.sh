#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): sharded-vs-single parity, strong-scaling bench lines for C2 and C3,
# and the C++ command-line program with one process per GPU.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_multi.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -s > gpurun_out/pytest_multi_gpu.log 2>&1; echo "multi rc=$?" >> gpurun_out/pytest_multi_gpu.log
grep -n "^\[\|passed\|failed\|skipped\|rc=" gpurun_out/pytest_multi_gpu.log | tail
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for cfg in c2 c3; do
  timeout 900 python bench.py --config $cfg --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --recon-iters 0 > gpurun_out/scale_${cfg}_n1.json 2> gpurun_out/scale_${cfg}_n1.err; echo "$cfg n1 rc=$?"
  timeout 900 $TR --master-port 29551 bench.py --config $cfg --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --recon-iters 0 > gpurun_out/scale_${cfg}_n$N.json 2> gpurun_out/scale_${cfg}_n$N.err; echo "$cfg n$N rc=$?"
  grep -h '^{' gpurun_out/scale_${cfg}_n1.json gpurun_out/scale_${cfg}_n$N.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n_gpus'], 'GPUs', round(d['value'], 2), d['unit'], round(d['ms_per_step'], 2), 'ms/step', 'e2e', round(d['e2e']['value'], 2), 'collectives', d.get('collectives'), d['config']['sharding'])
"
done
# the C++ program, one process per GPU (torchrun only exports RANK/WORLD_SIZE/LOCAL_RANK)
python -c "
import sys; sys.path.insert(0, '.')
from gpuvmem_b200 import synth
synth.write_gvms(synth.make_problem(N=512, nvis=200000, nchan=4, freq0=1.0e11, bandwidth=4e9, seed=5), '/tmp/cli.gvms')
"
mkdir -p /tmp/cli_out
rm -f /tmp/gvm_nccl_29552.id
GVM_OPTIMIZER=CG-FRPRMN timeout 600 $TR --master-port 29552 --no-python gpuvmem_b200/bin/gpuvmem -i /tmp/cli.gvms -m /tmp/cli.gvms -o /tmp/cli_out/res.gvmr -O /tmp/cli_out/img.f32 -p /tmp/cli_out/ -z 0.001,0.0 -Z 0.01 -t 5 -G $(seq -s, 0 $((N-1))) -f /tmp/cli_out/stats.txt > gpurun_out/cli_n$N.log 2>&1; echo "cli n$N rc=$?"
timeout 600 gpuvmem_b200/bin/gpuvmem -i /tmp/cli.gvms -m /tmp/cli.gvms -o /tmp/cli_out/res1.gvmr -O /tmp/cli_out/img1.f32 -p /tmp/cli_out/ -z 0.001,0.0 -Z 0.01 -t 5 -f /tmp/cli_out/stats1.txt > gpurun_out/cli_n1.log 2>&1; echo "cli n1 rc=$?"
ls -la /tmp/cli_out | tail -n 8; cat /tmp/cli_out/stats.txt; echo; cat /tmp/cli_out/stats1.txt; echo
python -c "
import numpy as np
a = np.fromfile('/tmp/cli_out/img.f32', np.float32); b = np.fromfile('/tmp/cli_out/img1.f32', np.float32)
print('cli image N-GPU vs 1-GPU rel-L2', np.linalg.norm(a - b) / np.linalg.norm(b), a.shape)
"
