# round 2, GPU call E: tests with the two-level kernel, cfg5rgb A/B (two-level kernel vs the separate instance pass), the default bench (cfg4) and its reference arm
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --workload cfg5rgb --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
echo "== cfg5rgb two-level kernel"; timeout 600 $B 2>&1 | tail -1 | tee gpurun_out/bench_cfg5rgb_e_trace2.json | cut -c1-160
echo "== cfg5rgb separate pass, refill at 8"; B200PT_NO_TRACE2=1 timeout 600 $B 2>&1 | tail -1 | tee gpurun_out/bench_cfg5rgb_e_pass8.json | cut -c1-160
echo "== cfg5rgb separate pass, no refill"; B200PT_NO_TRACE2=1 B200PT_SPHERE_REFILL=1 timeout 600 $B 2>&1 | tail -1 | tee gpurun_out/bench_cfg5rgb_e_pass1.json | cut -c1-160
echo "== default bench"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_default_e.json | cut -c1-400
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference_e.json | cut -c1-300
