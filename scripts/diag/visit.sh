bash scripts/gpu_run.sh tests
bash scripts/gpu_run.sh bench
bash scripts/gpu_run.sh "ncu:k_grad_umma:c2"
bash scripts/gpu_run.sh "launches:c2"
