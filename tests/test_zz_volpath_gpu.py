"""GPU parity tests of VolPathIntegrator with a homogeneous medium around the scene (last item): the
images of the reference's `Integrator "volpath"` (tests/golden/render_volpath_*.pfm) bit for bit, and larger renders
against the oracle's VolPathLi.  Like the SampledSpectrum tests this path was written after the round's GPU time was
spent: collected last, first GPU run pending; tests/test_hostcheck.py replays the same functions against the CPU check
build of the sources."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits
from render_cases import EXTRA, RENDERS, VOLPATH, VOLPATH_BOUNDED
from test_gpu_parity import ctx  # noqa: F401  (module-scoped context fixture)

pytestmark = pytest.mark.gpu

STRATEGY = {"uniform": "LIGHTS_UNIFORM", "power": "LIGHTS_POWER", "spatial": "LIGHTS_SPATIAL"}


@pytest.mark.parametrize("gname", sorted(VOLPATH))
def test_volpath_render_vs_reference_pfm(pkg, abi, scenes, ob, ctx, gname):
    base, medium, strat = VOLPATH[gname]
    nt, mats, w, h, spp, depth, _, nl = RENDERS[base]
    ex = EXTRA.get(base, {})
    arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex.get("scene", {}))
    setup = scenes.RenderSetup(w, h, spp, max_depth=depth, strategy=getattr(abi, STRATEGY[strat]), integrator="volpath",
                               medium=medium, **ex.get("camera", {}))
    scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
    r = pkg.Render(scene, setup)
    r.render_tiles()
    rgb = r.read_rgb()
    ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % gname))
    nbad = int((bits(rgb) != bits(ref)).sum())
    if nbad:
        o = ob.Oracle(abi, arr)
        ob.set_volpath(o.lib, True, medium)
        ys, xs, _ = np.nonzero(bits(rgb) != bits(ref))
        y, x = int(ys[0]), int(xs[0])
        print("first differing pixel", x, y, rgb[y, x], ref[y, x])
        print("gpu samples", r.debug_pixel_samples(x, y))
        print("oracle samples", o.pixel_samples(setup, x, y))
        ob.set_volpath(o.lib, False)
    assert nbad == 0, "%d of %d components differ from the reference's volpath render" % (nbad, rgb.size)
    r.close()
    scene.close()


@pytest.mark.parametrize("mats,depth,strat,medium,kw", [
    (("matte", "glass", "metal", "plastic"), 8, "power", dict(sigma_a=(0.01, 0.02, 0.03), sigma_s=(0.4, 0.35, 0.3), g=0.6), {}),
    (("matte", "plastic"), 12, "spatial", dict(sigma_a=(0.2, 0.2, 0.2), sigma_s=(1.5, 1.2, 0.9), g=-0.5), {}),   # thick: long chains
    (("matte", "metal"), 5, "uniform", dict(sigma_a=(0.05, 0.05, 0.05), sigma_s=(0.1, 0.1, 0.1), g=0.0), {"pixel_filter": "gaussian"}),
    (("matte", "glass"), 6, "uniform", None, {}),                                                                    # volpath without a medium
    (("matte", "plastic"), 7, "spatial", dict(sigma_a=(0.01, 0.02, 0.03), sigma_s=(0.4, 0.35, 0.3), g=0.6),
     {"sampler": "halton", "lens_radius": 0.05, "focal_distance": 4.0})])                                            # Halton, thin lens
def test_volpath_render_vs_oracle(pkg, abi, scenes, ob, ctx, mats, depth, strat, medium, kw):
    """Larger renders (several batches of work per bounce, Russian roulette in the medium, the general film path) against
    the oracle's VolPathLi: raw film sums bit for bit; every ray the reference traces as a closest-hit query is counted."""
    arr = scenes.SceneArrays(20000, materials=mats, soup_version=1)
    setup = scenes.RenderSetup(64, 48, 8, max_depth=depth, strategy=getattr(abi, STRATEGY[strat]), integrator="volpath",
                               medium=medium, **kw)
    scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
    o = ob.Oracle(abi, arr)
    ob.set_volpath(o.lib, True, medium)
    try:
        film, ostats = o.render(setup)
    finally:
        ob.set_volpath(o.lib, False)
    r = pkg.Render(scene, setup)
    r.render_tiles()
    raw = r.read_raw()
    assert int((bits(raw) != bits(film)).sum()) == 0
    st = r.stats()
    assert st["camera_rays"] == ostats["camera_rays"]
    # VisibilityTester::Tr uses Scene::Intersect, so the reference (and the oracle) count shadow rays as regular ones
    assert st["regular_rays"] + st["shadow_rays"] == ostats["regular_rays"] + ostats["shadow_rays"]
    r.close()
    scene.close()
    o.close()


from test_dropin_plugin import PLUGIN, needs_plugin  # noqa: E402


@needs_plugin
def test_volpath_dropin_binary_matches_reference(scenes, tmp_path):
    """`Integrator "volpath"`, `MakeNamedMedium` / `MediumInterface` parsed by the reference's own code and flattened by
    gpupath.cpp (GpuIntegrator<VolPathIntegrator>): bit-identical to the reference's image."""
    import subprocess
    for gname in ("volpath_fog", "volpath_fog_spheres", "volpath_four"):
        base, medium, strat = VOLPATH[gname]
        nt, mats, w, h, spp, depth, _, nl = RENDERS[base]
        ex = EXTRA.get(base, {})
        arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex.get("scene", {}))
        path = scenes.write_pbrt(str(tmp_path), "render_" + gname, arr, w, h, spp, max_depth=depth, strategy=strat,
                                 integrator="volpath", medium=medium, **ex.get("camera", {}))
        r = subprocess.run([PLUGIN, "--quiet", os.path.basename(path)], cwd=str(tmp_path), capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        got = scenes.read_pfm(os.path.join(str(tmp_path), "render_%s.pfm" % gname))
        ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % gname))
        assert np.array_equal(bits(got), bits(ref)), "drop-in volpath render (%s) differs from the reference" % gname


@needs_plugin
def test_bounded_media_dropin_binary_matches_reference(scenes, tmp_path):
    """`MakeNamedMedium` + `MediumInterface "cloud" "fog"` + `Material ""` on spheres, parsed by the reference's own code:
    gpupath.cpp turns the null-material spheres into medium boundaries (b200pt_integrator_desc::bounded_media)."""
    import subprocess
    for gname, (fog, spheres) in sorted(VOLPATH_BOUNDED.items()):
        arr = scenes.SceneArrays(3000, materials=("matte", "glass", "metal", "plastic"), soup_version=1, spheres=spheres)
        path = scenes.write_pbrt(str(tmp_path), "render_" + gname, arr, 40, 32, 8, max_depth=6, strategy="uniform",
                                 integrator="volpath", medium=fog)
        r = subprocess.run([PLUGIN, "--quiet", os.path.basename(path)], cwd=str(tmp_path), capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "falling back" not in (r.stdout + r.stderr).lower(), r.stdout + r.stderr
        got = scenes.read_pfm(os.path.join(str(tmp_path), "render_%s.pfm" % gname))
        ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % gname))
        assert np.array_equal(bits(got), bits(ref)), "drop-in render with bounded media (%s) differs from the reference" % gname


def test_volpath_instances_and_partial_spheres_vs_oracle(pkg, abi, scenes, ob, ctx):
    """The medium pass re-derives the hit distance for every kind of hit: top-level triangles, triangles of instanced
    objects (ray taken into the object's space), full and partial spheres."""
    medium = dict(sigma_a=(0.03, 0.04, 0.05), sigma_s=(0.25, 0.2, 0.3), g=0.3)
    scene_kw = dict(EXTRA["instances"]["scene"])
    scene_kw["spheres"] = EXTRA["sphere_partial"]["scene"]["spheres"]
    arr = scenes.SceneArrays(2000, materials=("matte", "glass", "metal", "plastic"), soup_version=1, **scene_kw)
    setup = scenes.RenderSetup(48, 40, 8, max_depth=6, strategy=abi.LIGHTS_POWER, integrator="volpath", medium=medium)
    scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
    o = ob.Oracle(abi, arr)
    ob.set_volpath(o.lib, True, medium)
    try:
        film, _ = o.render(setup)
    finally:
        ob.set_volpath(o.lib, False)
    r = pkg.Render(scene, setup)
    r.render_tiles()
    assert int((bits(r.read_raw()) != bits(film)).sum()) == 0
    r.close()
    scene.close()
    o.close()


@pytest.mark.parametrize("name", ["analytic_point", "analytic_4points", "analytic_area"])
def test_analytic_scenes_known_answer_volpath(pkg, abi, scenes, ob, ctx, name):
    """The reference's own known-answer test also runs VolPathIntegrator (depth 8) over its analytic scenes
    (src/tests/analytic_scenes.cpp:326-351, CheckSceneAverage :54-66): the image mean is 1 +- 0.02.  Here the image must
    in addition equal the oracle's VolPathLi bit for bit."""
    nt, mats, w, h, spp, depth, strat, nl = RENDERS[name]
    ex = EXTRA[name]
    arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex["scene"])
    setup = scenes.RenderSetup(w, h, spp, max_depth=depth, strategy=abi.LIGHTS_SPATIAL, integrator="volpath", **ex["camera"])
    scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
    r = pkg.Render(scene, setup)
    r.render_tiles()
    rgb = r.read_rgb()
    assert abs(float(rgb.mean()) - 1.0) < 0.02
    o = ob.Oracle(abi, arr)
    ob.set_volpath(o.lib, True, None)
    try:
        film, _ = o.render(setup, threads=4)
    finally:
        ob.set_volpath(o.lib, False)
    assert np.array_equal(bits(rgb), bits(o.film_rgb(setup, film)))
    r.close()
    scene.close()
    o.close()


@pytest.mark.parametrize("gname", sorted(VOLPATH_BOUNDED))
def test_volpath_bounded_media_vs_reference_pfm(pkg, abi, scenes, ob, ctx, gname):
    """Media bounded by null-material spheres (medium transitions): a cloud in vacuum; two clouds and a
    sphere light inside a fog.  The path steps over the boundaries without spending a bounce (volpath.cpp:115-121), the
    ray's medium switches by the side it leaves on (interaction.h:80-82), shadow and MIS rays collect transmittance
    segment by segment (light.cpp:63-81, scene.cpp:57-70): the reference's image bit for bit, and the ray counters of
    the oracle (every segment is a Scene::Intersect call)."""
    fog, spheres = VOLPATH_BOUNDED[gname]
    arr = scenes.SceneArrays(3000, materials=("matte", "glass", "metal", "plastic"), soup_version=1, spheres=spheres)
    setup = scenes.RenderSetup(40, 32, 8, max_depth=6, strategy=abi.LIGHTS_UNIFORM, integrator="volpath", medium=fog,
                               boundaries=arr.sphere_specs)
    scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
    r = pkg.Render(scene, setup)
    r.render_tiles()
    rgb = r.read_rgb()
    ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % gname))
    o = ob.Oracle(abi, arr)
    ob.set_volpath(o.lib, True, fog)
    ob.set_medium_boundaries(o.lib, arr)
    try:
        film, ostats = o.render(setup, threads=4)
        nbad = int((bits(rgb) != bits(ref)).sum())
        if nbad:
            ys, xs, _ = np.nonzero(bits(rgb) != bits(ref))
            y, x = int(ys[0]), int(xs[0])
            print("first differing pixel", x, y, rgb[y, x], ref[y, x])
            print("gpu samples", r.debug_pixel_samples(x, y))
            print("oracle samples", o.pixel_samples(setup, x, y))
    finally:
        ob.set_volpath(o.lib, False)  # also forgets the boundaries
    assert nbad == 0, "%d of %d components differ from the reference's render" % (nbad, rgb.size)
    assert int((bits(r.read_raw()) != bits(film)).sum()) == 0
    st = r.stats()
    assert st["camera_rays"] == ostats["camera_rays"]
    assert st["regular_rays"] + st["shadow_rays"] == ostats["regular_rays"] + ostats["shadow_rays"]
    assert st["dimension_overflows"] == 0 and st["stack_overflows"] == 0
    r.close()
    scene.close()
    o.close()


def test_bounded_media_larger_render_vs_oracle(pkg, abi, scenes, ob, ctx):
    """More paths than one warp-batch per pass, nested passes (a cloud behind a cloud), power light sampling, a thick
    and a thin cloud, a glass sphere between them: raw film sums and ray counters against the oracle's VolPathLi."""
    thick = dict(sigma_a=(0.5, 0.4, 0.3), sigma_s=(4.0, 5.0, 6.0), g=0.7)
    thin = dict(sigma_a=(0.05, 0.05, 0.1), sigma_s=(0.3, 0.3, 0.2), g=-0.3)
    spheres = (dict(center=(0.3, 0.1, -2.4), radius=0.6, boundary=thick), dict(center=(-0.5, -0.2, -1.2), radius=0.8, boundary=thin),
               dict(center=(0.9, 0.9, -1.9), radius=0.3, material="glass"), dict(center=(-1.3, 1.6, -1.6), radius=0.3, emit=80.0))
    for fog in (None, dict(sigma_a=(0.03, 0.03, 0.03), sigma_s=(0.15, 0.1, 0.1), g=0.1)):
        arr = scenes.SceneArrays(20000, materials=("matte", "glass", "metal", "plastic"), soup_version=1, spheres=spheres)
        setup = scenes.RenderSetup(64, 48, 8, max_depth=9, strategy=abi.LIGHTS_POWER, integrator="volpath", medium=fog,
                                   boundaries=arr.sphere_specs)
        scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
        o = ob.Oracle(abi, arr)
        ob.set_volpath(o.lib, True, fog)
        ob.set_medium_boundaries(o.lib, arr)
        try:
            film, ostats = o.render(setup, threads=4)
        finally:
            ob.set_volpath(o.lib, False)
        r = pkg.Render(scene, setup)
        r.render_tiles()
        assert int((bits(r.read_raw()) != bits(film)).sum()) == 0
        st = r.stats()
        assert st["regular_rays"] + st["shadow_rays"] == ostats["regular_rays"] + ostats["shadow_rays"]
        assert st["dimension_overflows"] == 0
        r.close()
        scene.close()
        o.close()
