"""GPU parity tests of feature combinations (collected last, first GPU run pending like the other test_zz files; replayed
against the CPU check build by tests/test_hostcheck.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits
from render_cases import EXTRA, RENDERS
from test_gpu_parity import ctx  # noqa: F401  (module-scoped context fixture)

pytestmark = pytest.mark.gpu
STRATEGY = {"uniform": "LIGHTS_UNIFORM", "power": "LIGHTS_POWER", "spatial": "LIGHTS_SPATIAL"}


COMBOS = [
    # spectral, integrator, medium, sampler, pixel filter, strategy, scene extras, materials
    (True, "path", None, "halton", "mitchell", "spatial", "spheres", ("matte", "glass", "metal", "plastic")),
    (True, "volpath", "thin", "sobol", None, "power", "instances", ("matte", "glass", "metal", "plastic")),
    (False, "volpath", "thick", "halton", "gaussian", "spatial", "delta_lights", ("matte", "plastic")),
    (False, "volpath", "thin", "halton", None, "uniform", "sphere_partial", ("matte", "mirror", "glass", "plastic")),
    (False, "volpath", None, "sobol", "sinc", "power", "instances", ("matte_rough", "glass", "metal", "plastic")),
    (True, "path", None, "sobol", None, "power", "delta_lights", ("matte", "glass", "metal", "plastic")),
]


@pytest.mark.parametrize("combo", range(len(COMBOS)))
def test_feature_combinations_vs_oracle(pkg, abi, scenes, ob, ctx, combo):
    """The widenings were each pinned on their own scenes; here they meet: spectrum type x integrator x medium x sampler x
    pixel filter x light distribution x shape / light kinds, device (check build) against the oracle, raw film sums."""
    import json
    spectral, integrator, med, sampler, pfilter, strat, extra, mats = COMBOS[combo]
    media = {"thin": dict(sigma_a=(0.05, 0.08, 0.12), sigma_s=(0.3, 0.25, 0.2), g=0.4),
             "thick": dict(sigma_a=(0.1, 0.05, 0.02), sigma_s=(0.15, 0.2, 0.3), g=0.0)}
    medium = media.get(med)
    tables = json.load(open(os.path.join(GOLDEN, "spectral_tables.json"))) if spectral else None
    nl = RENDERS[extra][7]
    arr = scenes.SceneArrays(1500, materials=mats, soup_version=1, n_lights=nl, seed=100 + combo, **EXTRA[extra]["scene"])
    if spectral:
        arr.attach_spectral(tables)
    kw = dict(sampler=sampler, integrator=integrator, medium=medium)
    if pfilter:
        kw["pixel_filter"] = pfilter
    if spectral and medium:
        kw["spectral_tables"] = tables
    setup = scenes.RenderSetup(40, 24, 4 if sampler == "sobol" else 3, max_depth=6,
                               strategy=getattr(abi, STRATEGY[strat]), **kw)
    o = ob.Oracle(abi, arr, spectral_tables=tables)
    ob.set_volpath(o.lib, integrator == "volpath", medium)
    try:
        film, _ = o.render(setup, threads=4)
    finally:
        ob.set_volpath(o.lib, False)
    scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
    r = pkg.Render(scene, setup)
    r.render_tiles()
    raw = r.read_raw()
    assert np.array_equal(bits(raw), bits(film))
    r.close()
    scene.close()
    o.close()
