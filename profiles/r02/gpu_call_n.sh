# round 2, GPU call N: k_trace2 with the three space changes sharing one converged enter_space
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg5s cfg5rgb; do
timeout 900 python bench.py --workload $w --steps 2 --warmup 1 --e2e-steps 1 2>&1 | tail -1 | tee gpurun_out/bench_${w}_n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['kernel_ms_per_step'], d['parity'])"
done
M="smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum"
timeout 300 ncu --metrics $M -k regex:k_trace2 -s 3 -c 1 python profiles/profile_trace.py cfg5s 2>&1 | grep -E "k_trace|inst_executed|time_duration"
