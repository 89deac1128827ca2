"""CPU tests: the oracle restatement against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py) -- bit-exact everywhere."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits, golden_camera, hexf

from render_cases import EXTRA, RENDERS  # noqa: E402  (the table shared with tests/golden/make_golden.py)


def test_sobol_stream_matches_reference(abi, scenes, ob, probe_json):
    lib = ob.load(abi)
    for rec in probe_json["sobol"]:
        b = rec["bounds"]
        setup = scenes.RenderSetup(b[2], b[3], rec["spp"], camera=abi.CameraDesc())
        n = len(rec["values"])
        out = np.zeros(n, np.float32)
        import ctypes as C
        lib.oracle_sobol(C.byref(setup.sampler), rec["px"], rec["py"], rec["sample"], rec["dim0"], n, abi.ptr(out))
        assert np.array_equal(bits(out), bits(hexf(rec["values"])))


def test_halton_tables_and_stream_match_reference(abi, scenes, ob):
    """HaltonSampler: PCG32-shuffled digit permutations and the sample stream (incl. a 1x1 film, stride 1)."""
    import ctypes as C
    import json
    import os
    lib = ob.load(abi)
    tables = scenes.HaltonTables()
    lib.oracle_halton_permutations.restype = C.c_int64
    lib.oracle_halton_permutations.argtypes = [C.c_int32, C.c_void_p]
    n = lib.oracle_halton_permutations(tables.n_dims, None)
    assert n == tables.perms.size
    mine = np.zeros(n, np.uint16)
    lib.oracle_halton_permutations(tables.n_dims, abi.ptr(mine))
    assert np.array_equal(mine, tables.perms)
    for rec in json.load(open(os.path.join(scenes.GOLDEN_DIR, "probe_halton.json"))):
        b = rec["bounds"]
        setup = scenes.RenderSetup(b[2], b[3], rec["spp"], camera=abi.CameraDesc(), sampler="halton")
        setup.sampler.sample_bounds[:] = b
        nv = len(rec["values"])
        out = np.zeros(nv, np.float32)
        lib.oracle_sobol(C.byref(setup.sampler), rec["px"], rec["py"], rec["sample"], rec["dim0"], nv, abi.ptr(out))
        assert np.array_equal(bits(out), bits(hexf(rec["values"])))


def test_camera_rays_match_reference(abi, scenes, ob, probe_json):
    lib = ob.load(abi)
    import ctypes as C
    for rec in probe_json["camrays"]:
        w, h = rec["res"]
        cam = golden_camera(abi, probe_json, w, h)
        setup = scenes.RenderSetup(w, h, rec["spp"], camera=cam)
        n = len(rec["rays"])
        out = np.zeros(n, dtype=abi.RAY_DTYPE)
        lib.oracle_camera_rays(C.byref(cam), C.byref(setup.sampler), rec["px"], rec["py"], n, abi.ptr(out))
        want = np.array([hexf(r) for r in rec["rays"]])
        assert np.array_equal(bits(out["o"]), bits(want[:, 0:3]))
        assert np.array_equal(bits(out["d"]), bits(want[:, 3:6]))
        assert np.array_equal(bits(out["t_max"]), bits(want[:, 6]))


def test_host_camera_matches_reference(pkg, abi, probe_json):
    for key, rec in probe_json["cameras"].items():
        w, h = map(int, key.split("x"))
        cam = pkg.host_perspective_camera((0, 0, -4.5), (0, 0, 0), (0, 1, 0), 35.0, w, h)
        assert np.array_equal(bits(np.array(cam.raster_to_camera[:])), bits(hexf(rec["raster_to_camera"]))), key
        assert np.array_equal(bits(np.array(cam.camera_to_world[:])), bits(hexf(rec["camera_to_world"]))), key


def test_host_roughness_and_copper_constants(pkg, scenes, probe_json):
    for r, a in probe_json["consts"]["roughness_to_alpha"]:
        got = np.float32(pkg.host_roughness_to_alpha(float.fromhex(r)))
        assert bits(got) == bits(hexf([a]))[0]
    assert np.array_equal(bits(np.array(scenes.COPPER_ETA)), bits(hexf(probe_json["consts"]["copper_eta"][0])))
    assert np.array_equal(bits(np.array(scenes.COPPER_K)), bits(hexf(probe_json["consts"]["copper_k"][0])))


def test_intersections_match_reference_bvhaccel(abi, scenes, ob):
    rays = np.load(os.path.join(GOLDEN, "isect_rays.npy"))
    ref = np.load(os.path.join(GOLDEN, "isect_ref.npy"))
    arr = scenes.SceneArrays(3000, materials=("matte",), soup_version=1, seed=99)
    o = ob.Oracle(abi, arr)
    for brute in (False, True):
        hits = o.trace_closest(rays, brute=brute)
        assert np.array_equal(hits["triangle"], ref["tri"])
        m = ref["tri"] >= 0
        assert m.sum() > 500
        assert np.array_equal(bits(hits["t"][m]), bits(ref["t"][m]))
    occ = o.trace_any(rays)
    assert np.array_equal(occ.astype(np.int32), ref["occluded"])
    o.close()


def test_triangle_badcase_misses(abi, scenes, ob):
    # reference tests/shapes.cpp:544-559 (Triangle.BadCases): this ray must NOT hit this triangle
    arr = scenes.SceneArrays(1, materials=("matte",), soup_version=0)
    arr.vertices = np.array([[[-1113.45459, -79.049614, -56.2431908], [-1113.45459, -87.0922699, -56.2431908],
                              [-1113.45459, -79.2090149, -56.2431908]]], np.float32)
    arr.material_id = np.zeros(1, np.int32)
    arr.light_id = np.full(1, -1, np.int32)
    arr.flip = np.zeros(1, np.uint8)
    arr.n_lights = 0
    arr.ply_parts = []
    rays = np.zeros(1, dtype=abi.RAY_DTYPE)
    rays["o"] = [[-1081.47925, 99.9999542, 87.7701111]]
    rays["d"] = [[-32.1072998, -183.355865, -144.607635]]
    rays["t_max"] = 0.9999
    o = ob.Oracle(abi, arr)
    assert o.trace_closest(rays, brute=True)["triangle"][0] == -1
    assert o.trace_closest(rays)["triangle"][0] == -1
    o.close()


@pytest.mark.parametrize("name", sorted(RENDERS))
def test_render_matches_reference_pfm(abi, scenes, ob, probe_json, name):
    nt, mats, w, h, spp, depth, strat, nl = RENDERS[name]
    ex = EXTRA.get(name, {})
    arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex.get("scene", {}))
    setup = scenes.RenderSetup(w, h, spp, max_depth=depth,
                               strategy={"uniform": abi.LIGHTS_UNIFORM, "power": abi.LIGHTS_POWER, "spatial": abi.LIGHTS_SPATIAL}[strat],
                               **ex.get("camera", {}))
    o = ob.Oracle(abi, arr)
    film, stats = o.render(setup, threads=4)
    rgb = o.film_rgb(setup, film)
    ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % name))
    assert ref.shape == rgb.shape
    assert np.array_equal(bits(rgb), bits(ref)), "oracle render is not bit-identical to the reference PFM"
    cb = setup.sample_bounds
    assert stats["camera_rays"] == (cb[2] - cb[0]) * (cb[3] - cb[1]) * setup.sampler.samples_per_pixel
    o.close()


@pytest.mark.parametrize("gname,base", [("spectral_four", "four"), ("spectral_rough", "rough"), ("spectral_instances", "instances"),
                                        ("spectral_spheres", "spheres"), ("spectral_delta_lights", "delta_lights")])
def test_spectral_oracle_matches_sampled_spectrum_reference(abi, scenes, ob, gname, base):
    """SURVEY 8(f) row 3, second half (groundwork for the device path): the oracle compiled with 60 spectral bins
    (oracle/Makefile liboracle_spectral.so) against the reference compiled with `typedef SampledSpectrum Spectrum`
    (core/pbrt.h:124-125; oracle/Makefile ref_spectral) -- bit-identical images.  The CIE curves and the FromRGB /
    copper spectra come from that reference build (tests/golden/spectral_tables.json, make_golden.py)."""
    import json
    if not os.path.exists(ob.LIB_SPECTRAL_PATH):
        pytest.skip("oracle/liboracle_spectral.so not built")
    tables = json.load(open(os.path.join(GOLDEN, "spectral_tables.json")))
    nt, mats, w, h, spp, depth, strat, nl = RENDERS[base]
    ex = EXTRA.get(base, {})
    arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex.get("scene", {}))
    setup = scenes.RenderSetup(w, h, spp, max_depth=depth,
                               strategy={"uniform": abi.LIGHTS_UNIFORM, "power": abi.LIGHTS_POWER, "spatial": abi.LIGHTS_SPATIAL}[strat],
                               **ex.get("camera", {}))
    o = ob.Oracle(abi, arr, spectral_tables=tables)
    film, _ = o.render(setup, threads=4)
    rgb = o.film_rgb(setup, film)
    ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % gname))
    assert np.array_equal(bits(rgb), bits(ref)), "spectral oracle is not bit-identical to the SampledSpectrum reference"
    rgb_ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % base))
    assert not np.array_equal(bits(ref), bits(rgb_ref))  # the two Spectrum types really give different images
    o.close()
    # the same scene described the way a SampledSpectrum host describes it to the library (b200pt_scene_desc::
    # material_spectra / light_spectra / cie_xyz, filled by SceneArrays.attach_spectral): the oracle reads those tables
    arr2 = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex.get("scene", {})).attach_spectral(tables)
    assert arr2.desc().n_spectrum_samples == abi.SPECTRUM_SAMPLES
    o = ob.Oracle(abi, arr2)
    film, _ = o.render(setup, threads=4)
    assert np.array_equal(bits(o.film_rgb(setup, film)), bits(ref))
    o.close()


def test_volpath_oracle_matches_reference_volpath(abi, scenes, ob):
    """Groundwork for media (last item): the oracle's VolPathIntegrator::Li (volpath.cpp:60-188) with a
    homogeneous medium around the whole scene -- free-flight sampling, Henyey-Greenstein phase function, transmittance
    along shadow and MIS rays, light sampling at every vertex -- against the reference's images, bit for bit."""
    from render_cases import VOLPATH
    for gname, (base, medium, strat) in sorted(VOLPATH.items()):
        nt, mats, w, h, spp, depth, _, nl = RENDERS[base]
        ex = EXTRA.get(base, {})
        arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex.get("scene", {}))
        setup = scenes.RenderSetup(w, h, spp, max_depth=depth,
                                   strategy={"uniform": abi.LIGHTS_UNIFORM, "power": abi.LIGHTS_POWER, "spatial": abi.LIGHTS_SPATIAL}[strat],
                                   **ex.get("camera", {}))
        o = ob.Oracle(abi, arr)
        ob.set_volpath(o.lib, True, medium)
        try:
            film, _ = o.render(setup, threads=4)
            rgb = o.film_rgb(setup, film)
        finally:
            ob.set_volpath(o.lib, False)
        ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % gname))
        assert np.array_equal(bits(rgb), bits(ref)), gname
        o.close()
    # the 60-bin build of both (SampledSpectrum reference, `volpath`): the oracle compiled with ORACLE_NSPEC=60
    if os.path.exists(ob.LIB_SPECTRAL_PATH):
        import json
        tables = json.load(open(os.path.join(GOLDEN, "spectral_tables.json")))
        for gname, vname in (("spectral_volpath_fog", "volpath_fog"), ("spectral_volpath_fog_spheres", "volpath_fog_spheres")):
            base, medium, strat = VOLPATH[vname]
            nt, mats, w, h, spp, depth, _, nl = RENDERS[base]
            ex = EXTRA.get(base, {})
            arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex.get("scene", {}))
            setup = scenes.RenderSetup(w, h, spp, max_depth=depth,
                                       strategy={"uniform": abi.LIGHTS_UNIFORM, "power": abi.LIGHTS_POWER, "spatial": abi.LIGHTS_SPATIAL}[strat],
                                       **ex.get("camera", {}))
            o = ob.Oracle(abi, arr, spectral_tables=tables)
            ob.set_volpath(o.lib, True, medium)
            try:
                film, _ = o.render(setup, threads=4)
                rgb = o.film_rgb(setup, film)
            finally:
                ob.set_volpath(o.lib, False)
            assert np.array_equal(bits(rgb), bits(scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % gname)))), gname
            o.close()
    # without a medium VolPathIntegrator still is not PathIntegrator (it samples a light at purely specular vertices too)
    assert not np.array_equal(bits(scenes.read_pfm(os.path.join(GOLDEN, "render_volpath_four.pfm"))),
                              bits(scenes.read_pfm(os.path.join(GOLDEN, "render_four.pfm"))))


def test_volpath_oracle_with_bounded_media_matches_reference(abi, scenes, ob):
    """Groundwork for the next widening of media: null-material spheres around a homogeneous medium (a cloud in vacuum; two
    clouds and a sphere light inside a global fog).  The path skips the boundaries without spending a bounce
    (volpath.cpp:115-121), the ray's medium switches by the side it leaves on (interaction.h:80-82), shadow and MIS rays
    accumulate transmittance segment by segment (light.cpp:63-81, scene.cpp:57-70): bit-identical to the reference."""
    from render_cases import VOLPATH_BOUNDED
    for gname, (fog, spheres) in sorted(VOLPATH_BOUNDED.items()):
        arr = scenes.SceneArrays(3000, materials=("matte", "glass", "metal", "plastic"), soup_version=1, spheres=spheres)
        setup = scenes.RenderSetup(40, 32, 8, max_depth=6, strategy=abi.LIGHTS_UNIFORM)
        o = ob.Oracle(abi, arr)
        ob.set_volpath(o.lib, True, fog)
        ob.set_medium_boundaries(o.lib, arr)
        try:
            film, _ = o.render(setup, threads=4)
            rgb = o.film_rgb(setup, film)
        finally:
            ob.set_volpath(o.lib, False)
        assert np.array_equal(bits(rgb), bits(scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % gname)))), gname
        o.close()


@pytest.mark.parametrize("name", ["analytic_point", "analytic_4points", "analytic_area"])
def test_analytic_scenes_known_answer(abi, scenes, ob, name):
    """The reference's own known-answer test (src/tests/analytic_scenes.cpp:54-66, CheckSceneAverage): the mean of the
    rendered image is 1 +- 0.02 -- for the reference's image and for the oracle's."""
    nt, mats, w, h, spp, depth, strat, nl = RENDERS[name]
    ex = EXTRA[name]
    arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex["scene"])
    setup = scenes.RenderSetup(w, h, spp, max_depth=depth, strategy=abi.LIGHTS_SPATIAL, **ex["camera"])
    o = ob.Oracle(abi, arr)
    film, _ = o.render(setup, threads=4)
    rgb = o.film_rgb(setup, film)
    ref = scenes.read_pfm(os.path.join(GOLDEN, "render_%s.pfm" % name))
    assert abs(float(ref.mean()) - 1.0) < 0.02
    assert abs(float(rgb.mean()) - 1.0) < 0.02
    o.close()


def test_pixelbounds_shards_equal_full_render(abi, scenes, ob):
    # tile shards with the full-film sampler are bit-identical to the full render
    arr = scenes.SceneArrays(2000, materials=("matte", "plastic"), soup_version=1)
    setup = scenes.RenderSetup(48, 32, 4)
    o = ob.Oracle(abi, arr)
    full, _ = o.render(setup, threads=2)
    tiles = np.arange(setup.n_tiles)
    a, _ = o.render(setup, tiles=tiles[0::2], threads=2)
    b, _ = o.render(setup, tiles=tiles[1::2], threads=2)
    assert np.array_equal(bits(a + b), bits(full))
    o.close()
