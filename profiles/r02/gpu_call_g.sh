# round 2, GPU call G: instance entry / exit batched in the two-level kernel
python -m pytest tests -m gpu -x -q -k "instance or trace or render_vs_reference or combos or volpath" 2>&1 | tail -3
B="python bench.py --workload cfg5rgb --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
echo "== cfg5rgb two-level kernel, batched space changes"; timeout 600 $B 2>&1 | tail -1 | tee gpurun_out/bench_cfg5rgb_g.json | cut -c1-160
