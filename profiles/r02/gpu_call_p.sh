# round 2, GPU call P: 60-bin shading kernels, resident CTAs per SM vs the local-memory working set (6.4 KB per thread)
for c in 8 2 1; do
echo "== B200PT_S60_SHADE_CTAS=$c"
B200PT_S60_SHADE_CTAS=$c timeout 600 python bench.py --workload cfg5 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_p_ctas$c.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
done
