# ncu captures of the kernels added after the first profile round (run under gpurun; outputs <= 64 MiB in total)
set -x
export B200PT_BATCH_PATHS=4194304
NCU="ncu --set full --clock-control none --import-source on"
timeout 280 $NCU -k regex:k_spheres -s 3 -c 1 -f -o gpurun_out/prof_spheres_cfg5rgb python profiles/profile_trace.py cfg5rgb > gpurun_out/prof_a.log 2>&1
PROFILE_PIXEL_FILTER=gaussian timeout 200 $NCU -k regex:k_film_tile -c 1 -f -o gpurun_out/prof_film_gaussian python profiles/profile_trace.py cfg2 > gpurun_out/prof_b.log 2>&1
PROFILE_BVH=gpu timeout 200 $NCU -k 'regex:k_lbvh_(karras|fit|collapse)' -c 5 -f -o gpurun_out/prof_lbvh_cfg2 python profiles/profile_trace.py cfg2 > gpurun_out/prof_c.log 2>&1
unset B200PT_BATCH_PATHS
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_d.log 2>&1
ls -la gpurun_out/
du -sh gpurun_out
