"""Times a few wavefront batches of a workload (device time per kernel class) -- used to
sweep env-var knobs (B200PT_REFILL_LANES, B200PT_BATCH_PATHS, B200PT_SAH_CPRIM) on the GPU box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n_batches = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pkg = graft.load_package()
from pbrt_v3_distributed_b200 import scenes  # noqa: E402

n_tris, mats, xres, yres, spp, depth, n_lights, _ = bench.WORKLOADS[name]
arr = scenes.SceneArrays(n_tris, materials=mats, soup_version=1, n_lights=n_lights)
setup = scenes.RenderSetup(xres, yres, spp, max_depth=depth)
ctx = pkg.Context(0)
t0 = time.time()
scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
build_s = time.time() - t0
r = pkg.Render(scene, setup)
per_batch = max(1, int(os.environ.get("B200PT_BATCH_PATHS", 16 << 20)) // (256 * spp))
mid = (r.tiles_y // 2) * r.tiles_x + r.tiles_x // 4
tiles = (mid + np.arange(per_batch * n_batches)) % r.n_tiles
r.render_tiles(tiles[:per_batch])  # warm-up
ctx.synchronize()
r.reset_stats()
r.set_option("profile", 1)
t0 = time.time()
r.render_tiles(tiles)
ctx.synchronize()
wall = time.time() - t0
st = r.stats()
rays = st["regular_rays"] + st["shadow_rays"]
print("%s refill=%s batch=%s postpone=%s: build %.1fs; %.1f Mrays/s wall (%.1f ms); closest %.1f ms, any %.1f ms, other %.1f ms; "
      "closest-only %.1f Mrays/s" % (name, os.environ.get("B200PT_REFILL_LANES", "-"), os.environ.get("B200PT_BATCH_PATHS", "-"),
                                    os.environ.get("B200PT_POSTPONE_PCT", "-"), build_s, rays / wall / 1e6, wall * 1e3,
                                    st["closest_ms"], st["any_ms"], st["shade_ms"],
                                    st["regular_rays"] / max(st["closest_ms"], 1e-9) / 1e3))
