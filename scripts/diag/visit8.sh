nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --config c4l --steps 3 --warmup 3 --no-cpu-baseline --no-configs --recon-iters 2 > gpurun_out/c4l_n8.json 2> gpurun_out/c4l_n8.err; echo "c4l n8 rc=$?"
tail -c 2500 gpurun_out/c4l_n8.json; tail -3 gpurun_out/c4l_n8.err
