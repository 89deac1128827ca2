# round 2, GPU call H (2 GPUs): bench.py under torchrun with the library's own film reduce; the drop-in binary as two ranks with NCCL
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_default_2gpu_h.json | cut -c1-300
# drop-in: two pbrt_b200 processes, one per GPU, film merged by b200pt_film_reduce (NCCL id through a file)
python - <<'PY'
import os, subprocess, sys, tempfile
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from pbrt_v3_distributed_b200 import scenes
import test_dropin_plugin as P
tmp = tempfile.mkdtemp()
path = P._scene(scenes, tmp)
procs = []
for k in range(2):
    env = dict(os.environ, B200PT_RANK=str(k), B200PT_WORLD_SIZE="2", B200PT_DEVICE=str(k), B200PT_NCCL_ID_FILE=os.path.join(tmp, "nccl.id"))
    procs.append(subprocess.Popen([P.PLUGIN, "--quiet", os.path.basename(path)], cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
logs = [p.communicate(timeout=300)[0] for p in procs]
print("return codes", [p.returncode for p in procs]); print("\n".join(l[-400:] for l in logs))
got = scenes.read_pfm(os.path.join(tmp, "render_four.pfm")); ref = scenes.read_pfm(os.path.join(P.GOLDEN, "render_four.pfm"))
print("two ranks on two GPUs, NCCL film reduce: image bit-identical to the reference:", bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32))), "leftover rank files:", [f for f in os.listdir(tmp) if ".rank" in f])
PY
