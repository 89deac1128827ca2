# round 2, GPU call C: tests, CTAs-per-SM sweep, child-prefetch variant, launch list, ncu captures (plain vs TMA-staged top)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
S="trace_ctas=6 trace_ctas=7 trace_ctas=8 trace_ctas=8,overlap=0"
timeout 300 python profiles/sweep2.py cfg3 3 $S 2>&1 | tee gpurun_out/sweep2_cfg3_c.log
B200PT_LIB_VARIANT=prefetch timeout 300 python profiles/sweep2.py cfg3 3 trace_ctas=7 trace_ctas=8 2>&1 | tee gpurun_out/sweep2_cfg3_c_prefetch.log
export B200PT_TRACE_CTAS_RT=8
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2c_cfg3.csv python profiles/profile_trace.py cfg3 > gpurun_out/prof_r2c_l.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_r2c_trace_cfg3_c8 python profiles/profile_trace.py cfg3 > gpurun_out/prof_r2c.log 2>&1
B200PT_STAGE_NODES=57 timeout 300 $NCU -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/prof_r2c_trace_cfg3_c8_tma57 python profiles/profile_trace.py cfg3 > gpurun_out/prof_r2c_tma.log 2>&1
ls -la gpurun_out/*r2c*
