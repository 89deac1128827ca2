# round 2, GPU call T: final state (lobe-kind masks + BSDF entry points outlined; 60-bin build with lazy spectra and 16 Mi-slot
# batches): GPU tests, the default bench, cfg5 / cfg5rgb, launch list of one cfg4 batch, ncu of the shading kernels
# (summarised on the box: the reports are too large to bring back)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 6 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_default_t.json | cut -c1-400
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --e2e-steps 1 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_t.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['parity'])"
timeout 600 python bench.py --workload cfg5rgb --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg5rgb_t.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2t_cfg4.csv python profiles/profile_trace.py cfg4 > gpurun_out/prof_r2t_l.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:k_shade -s 4 -c 4 -f -o /tmp/prof_r2t_shade_cfg4 python profiles/profile_trace.py cfg4 > gpurun_out/prof_r2t_shade.log 2>&1
python profiles/ncu_summary.py /tmp/prof_r2t_shade_cfg4.ncu-rep "k_shade<matte|plastic|metal|glass, lean>, the four bounce-1 shading launches of one 16 Mi-slot cfg4 batch (profiles/profile_trace.py cfg4), final build: lobe-kind masks + outlined BSDF entry points" > gpurun_out/ncu_shade_cfg4_after.json
timeout 400 $NCU -k regex:k_shade -s 4 -c 4 -f -o /tmp/prof_r2t_shade_cfg5 python profiles/profile_trace.py cfg5 > gpurun_out/prof_r2t_shade5.log 2>&1
python profiles/ncu_summary.py /tmp/prof_r2t_shade_cfg5.ncu-rep "k_shade<.., general> of the 60-bin build (lazy spectra), the four bounce-1 shading launches of one 16 Mi-slot cfg5 batch (profiles/profile_trace.py cfg5)" > gpurun_out/ncu_shade_cfg5_lazy.json
ncu -i /tmp/prof_r2t_shade_cfg4.ncu-rep --page source --csv -k regex:"k_shade<1" 2>/dev/null | head -c 6000000 > gpurun_out/shade_plastic_source.csv
ls -la gpurun_out/
