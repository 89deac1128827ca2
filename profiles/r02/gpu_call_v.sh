# round 2, GPU call V (2 GPUs): the bench workload under torchrun with the library's film reduce; per-rank render time of the
# static tile split (tile i -> rank i mod N) in the JSON line
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_default_2gpu_v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['n_gpus'], d['per_rank_render_ms_per_step'], d['e2e'])"
