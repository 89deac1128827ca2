"""BASELINE.json configs[0] (killeroo-simple.pbrt at 400x400, 64 spp) through the drop-in binary vs the unmodified
reference, both reading the same .pbrt file (staged by `make -C oracle ref`).  Prints one JSON line.

  python profiles/killeroo_cfg1.py            # on a B200 box
"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "oracle", "_ref", "scenes")
PLUGIN = os.path.join(ROOT, "pbrt-v3-distributed_b200", "_plugin", "pbrt_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "pbrt_ref")


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "killeroo-cfg1.pbrt"
    tmp = tempfile.mkdtemp()
    rep = os.path.join(tmp, "report.json")
    env = dict(os.environ, B200PT_REPORT=rep, B200PT_REPEAT="4")
    t0 = time.time()
    subprocess.run([PLUGIN, "--quiet", "--outfile", os.path.join(tmp, "gpu.pfm"), scene], cwd=SCENES, env=env, check=True)
    wall_gpu4 = time.time() - t0
    env1 = dict(os.environ, B200PT_REPORT=rep + "1")
    t0 = time.time()
    subprocess.run([PLUGIN, "--quiet", "--outfile", os.path.join(tmp, "gpu1.pfm"), scene], cwd=SCENES, env=env1, check=True)
    wall_gpu = time.time() - t0
    r = json.load(open(rep))
    cold = json.load(open(rep + "1"))
    t0 = time.time()
    subprocess.run([REF, "--quiet", "--outfile", os.path.join(tmp, "ref.pfm"), scene], cwd=SCENES, check=True)
    wall_ref = time.time() - t0
    same = open(os.path.join(tmp, "gpu1.pfm"), "rb").read() == open(os.path.join(tmp, "ref.pfm"), "rb").read()
    rays = r["regular_rays"] + r["shadow_rays"]
    out = {"workload": scene, "triangles": r["triangles"], "spheres": r["spheres"],
           "gpu_render_ms_warm": r["render_ms"], "gpu_render_ms_cold": cold["render_ms"],
           "gpu_scene_build_upload_ms": cold["scene_build_upload_ms"], "gpu_ctx_ms": cold["ctx_ms"],
           "mrays_per_s_warm": rays / r["render_ms"] / 1e3, "msamples_per_s_warm": r["camera_rays"] / r["render_ms"] / 1e3,
           "pbrt_b200_process_wall_s": wall_gpu, "pbrt_ref_process_wall_s": wall_ref, "host_cores": os.cpu_count(),
           "wall_speedup": wall_ref / wall_gpu, "launches_per_render": r["launches"],
           "images_bit_identical": same, "wall_of_4_repeats_s": wall_gpu4}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
