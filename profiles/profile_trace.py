"""Short single-GPU run for Nsight Compute: one wavefront batch of a bench workload.

  ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 3 -o gpurun_out/prof \
      python profiles/profile_trace.py cfg2
k_trace launch order inside a batch: [0] bounce-0 closest hit (camera rays), [1] bounce-0 shadow rays
(any hit), [2] bounce-0 MIS rays, [3] bounce-1 closest hit (incoherent), [4] bounce-1 shadow, ...
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n_batches = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pkg = graft.load_package()
from pbrt_v3_distributed_b200 import scenes  # noqa: E402

n_tris, mats, xres, yres, spp, depth, n_lights, _ = bench.WORKLOADS[name]
arr = scenes.SceneArrays(n_tris, materials=mats, soup_version=1, n_lights=n_lights, **bench.workload_scene_kwargs(name))
if name in bench.SPECTRAL_WORKLOADS:
    arr.attach_spectral(bench.spectral_tables())  # 60-bin host: the SampledSpectrum kernels
# PROFILE_PIXEL_FILTER=gaussian profiles the general film path, PROFILE_BVH=gpu the on-device builder
setup = scenes.RenderSetup(xres, yres, spp, max_depth=depth, pixel_filter=os.environ.get("PROFILE_PIXEL_FILTER"))
ctx = pkg.Context(0)
if os.environ.get("PROFILE_BVH") == "gpu":
    ctx.set_option("gpu_bvh_build", 1)
scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
r = pkg.Render(scene, setup)
per_batch = max(1, int(os.environ.get('B200PT_BATCH_PATHS', 16 << 20)) // (256 * spp))
# tiles from the middle of the film (the soup, not the background)
mid = (r.tiles_y // 2) * r.tiles_x + r.tiles_x // 4
tiles = (mid + np.arange(per_batch * n_batches)) % r.n_tiles
r.render_tiles(tiles)
ctx.synchronize()
print(r.stats())
