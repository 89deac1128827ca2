# round 2, GPU call S: shading kernels with the family's lobe-kind mask (default build) and with the BSDF entry points
# outlined (kmol1 / kmol2): GPU tests, cfg4 sweep (3 batches; 'other' = shading + raygen + film), cfg5 (60 bins) with the
# outlined recipe evaluation and with 16 Mi-slot batches, ncu of the shading kernels (cfg4 RGB, cfg5 60-bin)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "" kmol1 kmol2; do
echo "-- variant '$v'"
B200PT_LIB_VARIANT=$v timeout 300 python profiles/sweep2.py cfg4 3 overlap=0 "" 2>&1 | tail -2 | tee -a gpurun_out/sweep2_cfg4_s_shade.log
done
echo "== cfg5 default"
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_s.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
echo "== cfg5 kmol1"
B200PT_LIB_VARIANT=kmol1 timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_s_kmol1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
echo "== cfg5 default, 16 Mi-slot batches"
B200PT_BATCH_PATHS=16777216 timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_s_16mi.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:k_shade -s 4 -c 4 -f -o gpurun_out/prof_r2s_shade_cfg4 python profiles/profile_trace.py cfg4 > gpurun_out/prof_r2s_shade.log 2>&1
timeout 400 $NCU -k regex:k_shade -s 4 -c 4 -f -o gpurun_out/prof_r2s_shade_cfg5 python profiles/profile_trace.py cfg5 > gpurun_out/prof_r2s_shade5.log 2>&1
tail -2 gpurun_out/prof_r2s_shade5.log | cut -c1-200
ls -la gpurun_out/
