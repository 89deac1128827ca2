# round 2, GPU call O (8 GPUs): the default bench under torchrun, film merge through b200pt_film_reduce
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_default_8gpu_o.json | cut -c1-600
