"""GPU parity tests of the SampledSpectrum path: hosts built with `typedef SampledSpectrum
Spectrum` (core/pbrt.h:124-125) hand the library 60-bin spectra; the results must equal the image of the reference
built that way (tests/golden/render_spectral_*.pfm) and the 60-bin oracle, bit for bit.  Kept in a file of their own,
collected last: this path was written after the round's GPU time was spent, so its first run on a GPU is the next
`-m gpu` run (tests/test_hostcheck.py replays the same functions against the CPU check build of the sources)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, bits
from render_cases import EXTRA, RENDERS
from test_dropin_plugin import PLUGIN_SPECTRAL, needs_spectral_plugin
from test_gpu_parity import ctx  # noqa: F401  (module-scoped context fixture)

pytestmark = pytest.mark.gpu


def _spectral_tables():
    import json
    return json.load(open(os.path.join(GOLDEN, "spectral_tables.json")))


@pytest.mark.parametrize("gname,base", [("spectral_four", "four"), ("spectral_rough", "rough"), ("spectral_instances", "instances"),
                                        ("spectral_spheres", "spheres"), ("spectral_delta_lights", "delta_lights")])
def test_spectral_render_vs_sampled_spectrum_reference(pkg, abi, scenes, ob, ctx, gname, base):
    """a SampledSpectrum host (60-bin spectra in the descriptor) -- the image of the reference compiled
    with `typedef SampledSpectrum Spectrum` (tests/golden/render_spectral_*.pfm), bit for bit."""
    nt, mats, w, h, spp, depth, strat, nl = RENDERS[base]
    ex = EXTRA.get(base, {})
    arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex.get("scene", {})).attach_spectral(_spectral_tables())
    setup = scenes.RenderSetup(w, h, spp, max_depth=depth,
                               strategy={"uniform": abi.LIGHTS_UNIFORM, "power": abi.LIGHTS_POWER, "spatial": abi.LIGHTS_SPATIAL}[strat],
                               **ex.get("camera", {}))
    scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
    r = pkg.Render(scene, setup)
    r.render_tiles()
    rgb = r.read_rgb()
    ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % gname))
    nbad = int((bits(rgb) != bits(ref)).sum())
    if nbad:
        o = ob.Oracle(abi, arr)
        ys, xs, _ = np.nonzero(bits(rgb) != bits(ref))
        y, x = int(ys[0]), int(xs[0])
        print("first differing pixel", x, y, rgb[y, x], ref[y, x])
        print("gpu samples", r.debug_pixel_samples(x, y))
        print("oracle samples", o.pixel_samples(setup, x, y))
    assert nbad == 0, "%d of %d components differ from the SampledSpectrum reference render" % (nbad, rgb.size)
    r.close()
    scene.close()


@pytest.mark.parametrize("mats,depth,strat,pfilter", [(("matte", "glass", "metal", "plastic"), 8, "power", None),
                                                     (("matte", "plastic"), 5, "spatial", None),
                                                     (("matte", "metal"), 5, "uniform", "gaussian")])
def test_spectral_render_and_counters_vs_oracle(pkg, abi, scenes, ob, ctx, mats, depth, strat, pfilter):
    """Larger SampledSpectrum renders against the 60-bin oracle: raw film sums and ray counters, every light
    distribution, the general pixel-filter path, several batches."""
    kw = {"pixel_filter": pfilter} if pfilter else {}
    arr = scenes.SceneArrays(20000, materials=mats, soup_version=1).attach_spectral(_spectral_tables())
    setup = scenes.RenderSetup(64, 48, 8, max_depth=depth,
                               strategy={"uniform": abi.LIGHTS_UNIFORM, "power": abi.LIGHTS_POWER, "spatial": abi.LIGHTS_SPATIAL}[strat], **kw)
    scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
    o = ob.Oracle(abi, arr)
    film, ostats = o.render(setup)
    r = pkg.Render(scene, setup)
    r.render_tiles()
    raw = r.read_raw()
    assert int((bits(raw) != bits(film)).sum()) == 0
    st = r.stats()
    for k in ("camera_rays", "regular_rays", "shadow_rays"):
        assert st[k] == ostats[k], k
    r.close()
    scene.close()
    o.close()


@needs_spectral_plugin
def test_spectral_dropin_binary_matches_sampled_spectrum_reference(scenes, tmp_path):
    arr = scenes.SceneArrays(3000, materials=("matte", "glass", "metal", "plastic"), soup_version=1)
    path = scenes.write_pbrt(str(tmp_path), "render_spectral_four", arr, 40, 32, 8, max_depth=5, strategy="uniform")
    r = subprocess.run([PLUGIN_SPECTRAL, "--quiet", os.path.basename(path)], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = scenes.read_pfm(os.path.join(str(tmp_path), "render_spectral_four.pfm"))
    ref = scenes.read_pfm(os.path.join(GOLDEN, "render_spectral_four.pfm"))
    assert np.array_equal(bits(got), bits(ref)), "spectral drop-in render differs from the SampledSpectrum reference's PFM"


@pytest.mark.parametrize("gname,vname", [("spectral_volpath_fog", "volpath_fog"), ("spectral_volpath_fog_spheres", "volpath_fog_spheres")])
def test_spectral_volpath_render_vs_sampled_spectrum_reference(pkg, abi, scenes, ob, ctx, gname, vname):
    """Both widenings together: VolPathIntegrator with a homogeneous medium under a SampledSpectrum host (60-channel
    free-flight sampling and transmittance) against the 60-bin reference's `volpath` image."""
    from render_cases import VOLPATH
    base, medium, strat = VOLPATH[vname]
    nt, mats, w, h, spp, depth, _, nl = RENDERS[base]
    ex = EXTRA.get(base, {})
    tables = _spectral_tables()
    arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex.get("scene", {})).attach_spectral(tables)
    setup = scenes.RenderSetup(w, h, spp, max_depth=depth,
                               strategy={"uniform": abi.LIGHTS_UNIFORM, "power": abi.LIGHTS_POWER, "spatial": abi.LIGHTS_SPATIAL}[strat],
                               integrator="volpath", medium=medium, spectral_tables=tables, **ex.get("camera", {}))
    scene = pkg.Scene(ctx, arr.desc(), keepalive=arr)
    r = pkg.Render(scene, setup)
    r.render_tiles()
    ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % gname))
    assert np.array_equal(bits(r.read_rgb()), bits(ref))
    r.close()
    scene.close()


@needs_spectral_plugin
def test_spectral_volpath_dropin_binary_matches_sampled_spectrum_reference(scenes, tmp_path):
    """`Integrator "volpath"` + `MakeNamedMedium` through the SampledSpectrum host: gpupath.cpp hands over the medium's
    60-bin sigma_a / sigma_s (b200pt_medium::spectra)."""
    from render_cases import VOLPATH
    gname, vname = "spectral_volpath_fog", "volpath_fog"
    base, medium, strat = VOLPATH[vname]
    nt, mats, w, h, spp, depth, _, nl = RENDERS[base]
    arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl)
    path = scenes.write_pbrt(str(tmp_path), "render_" + gname, arr, w, h, spp, max_depth=depth, strategy=strat,
                             integrator="volpath", medium=medium)
    r = subprocess.run([PLUGIN_SPECTRAL, "--quiet", os.path.basename(path)], cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = scenes.read_pfm(os.path.join(str(tmp_path), "render_%s.pfm" % gname))
    ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % gname))
    assert np.array_equal(bits(got), bits(ref)), "spectral volpath drop-in render differs from the SampledSpectrum reference's PFM"
