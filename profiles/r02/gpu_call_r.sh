# round 2, GPU call R: 60-bin shading with lazy spectra (frame 6.4 KB -> 1.1 KB): all GPU tests, cfg5 with parity, the
# register variants of the lazy kernel (3 / 2 resident CTAs per SM: 168 / 255 registers, 0 spills at 2); then the RGB shading
# kernels on the bench workload: ncu of the four bounce-1 launches and the occupancy variants (5 / 6 / 3 CTAs per SM)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --workload cfg5 --steps 2 --warmup 1 --e2e-steps 1 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_r.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['parity'], d['cpu_baseline'])"
for v in s60c3 s60c2; do
echo "== variant $v"
B200PT_LIB_VARIANT=$v timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_r_$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
done
echo "== cfg5rgb for the ratio"
timeout 600 python bench.py --workload cfg5rgb --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg5rgb_r.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
echo "== ncu k_shade cfg4"
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:k_shade -s 4 -c 4 -f -o gpurun_out/prof_r2r_shade_cfg4 python profiles/profile_trace.py cfg4 > gpurun_out/prof_r2r_shade.log 2>&1
tail -2 gpurun_out/prof_r2r_shade.log | cut -c1-200
echo "== shading occupancy variants (cfg4, 3 batches, no overlap: 'other' = shading + raygen + film)"
for v in "" sh5 sh6 sh3; do
echo "-- variant '$v'"
B200PT_LIB_VARIANT=$v timeout 300 python profiles/sweep2.py cfg4 3 overlap=0 "" 2>&1 | tail -2 | tee -a gpurun_out/sweep2_cfg4_r_shade.log
done
ls -la gpurun_out/
