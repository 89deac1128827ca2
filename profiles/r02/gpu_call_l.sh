# round 2, GPU call L: the instanced scene as SURVEY 8(d) words it (50 x 1M), with full-size parity against the reference
timeout 1200 python bench.py --workload cfg5s --steps 2 --warmup 1 --e2e-steps 1 2>&1 | tail -1 | tee gpurun_out/bench_cfg5s_l.json | cut -c1-400
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:k_trace2 -s 3 -c 1 -f -o gpurun_out/prof_r2l_trace2_cfg5s python profiles/profile_trace.py cfg5s > gpurun_out/prof_r2l.log 2>&1
tail -3 gpurun_out/prof_r2l.log
