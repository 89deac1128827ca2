# round 2, GPU call Q: lobe spectra by reference in the 60-bin build (frame 6.4 -> 4.5 KB): tests + cfg5
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --e2e-steps 1 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_q.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['parity'])"
