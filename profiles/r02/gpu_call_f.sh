# round 2, GPU call F: all GPU tests, reference arm, cfg3 / cfg5 (60-bin) quick benches, ncu of the two-level kernel on cfg5rgb
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== reference arm (small)"; timeout 300 python bench.py --impl reference --steps 1 --warmup 1 --workload small 2>&1 | tail -1 | cut -c1-200
echo "== cfg3"; timeout 600 python bench.py --workload cfg3 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg3_f.json | cut -c1-200
echo "== cfg5 (SampledSpectrum)"; timeout 900 python bench.py --workload cfg5 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_f.json | cut -c1-200
export B200PT_BATCH_PATHS=4194304
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trace2 -s 3 -c 1 -f -o gpurun_out/prof_r2f_trace2_cfg5rgb python profiles/profile_trace.py cfg5rgb > gpurun_out/prof_r2f.log 2>&1
ls -la gpurun_out/prof_r2f*
