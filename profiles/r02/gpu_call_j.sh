# round 2, GPU call J: final kernel on the bench workload (cfg4): ncu capture + DRAM traffic, coherence-sort A/B, default bench
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_r2j_trace_cfg4 python profiles/profile_trace.py cfg4 > gpurun_out/prof_r2j.log 2>&1
timeout 400 $NCU -k regex:k_trace -s 1 -c 1 -f -o gpurun_out/prof_r2j_any_cfg4 python profiles/profile_trace.py cfg4 > gpurun_out/prof_r2j_any.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2j_cfg4.csv python profiles/profile_trace.py cfg4 > gpurun_out/prof_r2j_l.log 2>&1
timeout 400 python profiles/sweep2.py cfg4 3 "" overlap=0 2>&1 | tee gpurun_out/sweep2_cfg4_j.log
B200PT_SORT_FROM=1 timeout 400 python profiles/sweep2.py cfg4 3 "" 2>&1 | tee gpurun_out/sweep2_cfg4_j_sort1.log
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default_j.json | cut -c1-300
ls -la gpurun_out/*r2j*
