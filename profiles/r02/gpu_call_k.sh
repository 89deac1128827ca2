# round 2, GPU call K: bounded media on the device (new tests), triangle-phase variants (two parked groups / warp-wide pair list)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
M="smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum"
for v in "" park2 share; do
  echo "== variant ${v:-main}"
  B200PT_LIB_VARIANT=$v timeout 300 python -m pytest tests -m gpu -x -q -k "trace or render_vs_reference or parity" 2>&1 | tail -2
  B200PT_LIB_VARIANT=$v timeout 400 python profiles/sweep2.py cfg4 3 "" postpone_pct=60 overlap=0 2>&1 | tail -3
  B200PT_LIB_VARIANT=$v timeout 300 ncu --metrics $M -k regex:k_trace -s 3 -c 1 python profiles/profile_trace.py cfg4 2>&1 | grep -E "k_trace|inst_executed|time_duration" 
done 2>&1 | tee gpurun_out/variants_k.log
