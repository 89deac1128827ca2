# round 2, GPU call D: tests with the refilling sphere/instance pass, shared-memory stack depth A/B, cfg5rgb (instances)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in "" ss8 ss4 ss0; do echo "== variant ${v:-main}"; B200PT_LIB_VARIANT=$v timeout 300 python profiles/sweep2.py cfg3 3 "" overlap=0 2>&1 | tail -2; done | tee gpurun_out/sweep2_cfg3_d.log
timeout 600 python bench.py --workload cfg5rgb --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg5rgb_d.json
