"""CPU test of the N>1 protocol (world_size 2, gloo): every rank builds the full-film sampler,
renders the tiles scenes.rank_tiles() gives it, the raw film sums are reduced to rank 0 and must
equal the single-process film bit for bit.  The renderer on this CPU-only box is the oracle (the
GPU path is checked the same way by bench.py --gpus N on the B200 box)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT


def test_two_rank_tile_shards_reduce_to_full_film(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent('''
        import os, sys
        import numpy as np
        import torch, torch.distributed as dist
        sys.path.insert(0, %r)
        import __graft_entry__ as g
        pkg = g.load_package(); ob = g.load_oracle()
        from pbrt_v3_distributed_b200 import abi, scenes
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        arr = scenes.SceneArrays(3000, materials=("matte", "glass", "metal", "plastic"), soup_version=1)
        setup = scenes.RenderSetup(72, 40, 4)
        o = ob.Oracle(abi, arr)
        film, stats = o.render(setup, tiles=scenes.rank_tiles(setup.n_tiles, rank, world), threads=2)
        t = torch.from_numpy(film)
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
        rays = torch.tensor([stats["regular_rays"] + stats["shadow_rays"], stats["camera_rays"]], dtype=torch.float64)
        dist.all_reduce(rays)
        if rank == 0:
            full, fstats = o.render(setup, threads=2)
            same = np.array_equal(t.numpy().view(np.uint32), full.view(np.uint32))
            ok = same and rays[0].item() == fstats["regular_rays"] + fstats["shadow_rays"] and rays[1].item() == 72 * 40 * 4
            print("MULTIRANK_OK" if ok else "MULTIRANK_MISMATCH", same, rays.tolist())
        dist.destroy_process_group()
    ''' % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert "MULTIRANK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_rank_tiles_partition(scenes):
    for n, w in [(1, 1), (7, 2), (8160, 8), (5, 8)]:
        parts = [scenes.rank_tiles(n, r, w) for r in range(w)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
