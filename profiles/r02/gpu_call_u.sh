# round 2, GPU call U: 60-bin spectra slot-major ([capacity][60], float4 access) instead of [bin][capacity]: GPU tests and
# cfg5; A/B of the shading kernels' software prefetch (variant pf: next work item fetched one iteration ahead)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --e2e-steps 1 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_u.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['parity'])"
for v in "" pf; do
echo "-- variant '$v'"
B200PT_LIB_VARIANT=$v timeout 300 python profiles/sweep2.py cfg4 3 overlap=0 "" 2>&1 | tail -2 | tee -a gpurun_out/sweep2_cfg4_u_prefetch.log
done
