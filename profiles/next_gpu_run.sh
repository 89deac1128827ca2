#!/bin/bash
# The first GPU call after this round, in the order that spends the least budget on the most uncertainty
# (everything below has only run in the CPU check build of the sources, tests/emu):
#   gpurun --timeout 1500 -- 'bash profiles/next_gpu_run.sh'
# 1. the parity tests of the 60-bin kernels, of the homogeneous-medium volpath and of their combinations (seconds)
# 2. the whole -m gpu suite (the general k_shade variants changed by 1.6 %; everything else is byte-identical SASS)
# 3. numbers: cfg5 (instancing + SampledSpectrum), cfg2fog (volpath), and cfg2 again as the control
# 4. launch lists of the two new workloads (shares only) and one --set full capture each of k_shade (60 bins) and k_medium
set -x
mkdir -p gpurun_out
python -m pytest tests/test_zz_spectral_gpu.py tests/test_zz_volpath_gpu.py tests/test_zz_combos_gpu.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/new_paths_tests.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/gpu_tests.txt
python bench.py --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --workload cfg5 --steps 2 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
python bench.py --workload cfg2fog > gpurun_out/bench_cfg2fog.json 2> gpurun_out/bench_cfg2fog.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg5.csv \
    python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg2fog.csv \
    python bench.py --workload cfg2fog --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 1 -o gpurun_out/ncu_shade_s60 \
    python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_medium -s 2 -c 1 -o gpurun_out/ncu_medium \
    python bench.py --workload cfg2fog --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
