# round 2, GPU call I: box parameter rescaled on a closest hit (one compare per child), predicated hit-mask OR, no spills at 64 registers
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/sweep2.py cfg3 3 "" overlap=0 2>&1 | tail -3 | tee gpurun_out/sweep2_cfg3_i.log
