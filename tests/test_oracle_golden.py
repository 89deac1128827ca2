"""CPU tests: the oracle restatement against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py) -- bit-exact everywhere."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits, golden_camera, hexf

RENDERS = {
    "matte": (3000, ("matte",), 40, 32, 8, 5, "uniform", None),
    "four": (3000, ("matte", "glass", "metal", "plastic"), 40, 32, 8, 5, "uniform", None),
    "power16": (3000, ("matte", "glass", "metal", "plastic"), 32, 32, 4, 16, "power", 16),
    # thin-lens camera, two-sided lights, ReverseOrientation on the glass and plastic meshes
    "lens_flip": (3000, ("matte", "glass", "metal", "plastic"), 36, 24, 8, 6, "uniform", None),
    # pbrt's default light sample strategy (SpatialLightDistribution), 10 and 16 lights
    "spatial": (3000, ("matte", "glass", "metal", "plastic"), 40, 32, 8, 5, "spatial", None),
    "spatial16": (3000, ("matte", "plastic"), 32, 32, 4, 8, "spatial", 16),
    # per-vertex shading normals (matte + metal meshes) and uvs (matte + plastic meshes), flipped plastic
    "normals_uv": (3000, ("matte", "glass", "metal", "plastic"), 40, 32, 8, 6, "spatial", None),
    # film crop window (sampler built from the cropped sample bounds), film scale, maxsampleluminance
    "crop": (3000, ("matte", "glass", "metal", "plastic"), 70, 50, 4, 5, "uniform", None),
    # non-default lobes of the four materials: OrenNayar matte (sigma 30), rough glass (microfacet reflection +
    # transmission)
    "rough": (3000, ("matte_rough", "glass_rough", "metal", "plastic"), 40, 32, 8, 8, "spatial", None),
    # HaltonSampler (pbrt's default sampler): non-power-of-two sample counts, cropped sample bounds
    "halton": (3000, ("matte", "glass", "metal", "plastic"), 40, 32, 6, 5, "spatial", None),
    "halton_crop": (3000, ("matte", "plastic"), 70, 50, 3, 5, "uniform", None),
    # Sphere shapes: two sphere area lights (one under a uniform scale) next to the 10 quad lights, a glass
    # sphere, a plastic ellipsoid under a handedness-swapping scale, a matte sphere with ReverseOrientation
    "spheres": (3000, ("matte", "glass", "metal", "plastic"), 48, 40, 8, 6, "spatial", None),
    # the only light is a sphere (the situation of scenes/killeroo-simple.pbrt), Halton sampler
    "sphere_light": (3000, ("matte", "plastic"), 40, 32, 6, 5, "spatial", 0),
    "sphere_power": (3000, ("matte", "glass_rough"), 40, 32, 4, 7, "power", 4),
    # partial spheres (zmin / zmax / phimax clipping: std::atan2, second root), one of them an area light
    "sphere_partial": (3000, ("matte", "plastic"), 48, 40, 8, 5, "spatial", 4),
    # object instancing (TransformedPrimitive): two objects, five instances (one at the identity, one mirrored)
    "instances": (2000, ("matte", "glass", "metal", "plastic"), 48, 40, 8, 6, "spatial", None),
    # MirrorMaterial (SpecularReflection + FresnelNoOp) next to glass: long specular chains
    "mirror": (3000, ("matte", "mirror", "glass", "plastic"), 40, 32, 8, 8, "spatial", None),
    # delta lights (point, spot, distant) next to the area lights: no MIS branch, Light::Power / Sample_Li per kind
    "delta_lights": (3000, ("matte", "glass", "metal", "plastic"), 40, 32, 8, 5, "spatial", 4),
    "delta_power": (3000, ("matte", "plastic"), 40, 32, 4, 5, "power", 0),
    # pixel filters wider than the box (Film::filterTable weights, samples outside the film, tile aprons of 2-4
    # pixels); the reference image is rendered with --nthreads 1 so that its tile merge order is defined
    "filter_gaussian": (3000, ("matte", "glass", "metal", "plastic"), 40, 36, 4, 5, "spatial", None),
    "filter_mitchell": (3000, ("matte", "plastic"), 37, 33, 4, 5, "uniform", None),
    "filter_sinc": (3000, ("matte", "plastic"), 40, 32, 3, 5, "uniform", None),
    "filter_aniso_crop": (3000, ("matte", "metal"), 70, 50, 4, 5, "uniform", None),
    "filter_triangle": (3000, ("matte", "glass"), 33, 35, 4, 5, "power", None),
    "filter_box1": (3000, ("matte", "plastic"), 40, 32, 4, 5, "uniform", None),
}
EXTRA = {"lens_flip": dict(scene=dict(two_sided=True, reverse_orientation=(1, 3)),
                           camera=dict(lens_radius=0.05, focal_distance=4.5)),
         "normals_uv": dict(scene=dict(shading_normals=(0, 2), uvs=(0, 3), reverse_orientation=(3,))),
         "crop": dict(camera=dict(crop_window=(0.21, 0.83, 0.1, 0.74), film_scale=2.0, max_sample_luminance=9.0)),
         "halton": dict(camera=dict(sampler="halton")),
         "halton_crop": dict(camera=dict(sampler="halton", crop_window=(0.21, 0.83, 0.1, 0.74))),
         "spheres": dict(scene=dict(spheres=(
             dict(center=(1.2, 1.8, -1.5), radius=0.35, emit=60.0),
             dict(center=(-2.0, 0.5, -2.5), radius=0.2, emit=90.0, scale=(1.5, 1.5, 1.5)),
             dict(center=(0.1, 0.0, -2.2), radius=0.45, material="glass"),
             dict(center=(-0.9, -0.6, -2.0), radius=0.4, material="plastic", scale=(1.3, 0.7, -1.1)),
             dict(center=(0.9, -0.7, -1.9), radius=0.3, material="matte", reverse_orientation=True)))),
         "sphere_partial": dict(scene=dict(spheres=(
             dict(center=(0.1, 0.1, -2.2), radius=0.6, material="plastic", zmin=-0.3, zmax=0.45, phimax=250.0),
             dict(center=(-1.0, -0.6, -2.0), radius=0.5, material="matte", scale=(1.2, 0.8, 1.0), phimax=200.0,
                  reverse_orientation=True),
             dict(center=(1.1, 0.6, -1.8), radius=0.4, emit=80.0, zmin=-0.1, two_sided=True)))),
         "instances": dict(scene=dict(
             objects=(dict(n_tris=400, seed=5, material="plastic", size=0.35), dict(n_tris=150, seed=9, material="glass", size=0.5)),
             instances=(dict(object=0, center=(0.0, 0.0, -2.4)), dict(object=0, center=(1.2, 0.8, -2.0), scale=(0.7, 1.4, 1.0)),
                        dict(object=1, center=(-1.1, -0.7, -2.2), scale=(1.0, 1.0, -1.3)), dict(object=1),
                        dict(object=0, center=(-1.3, 0.9, -1.9), scale=(1.5, 1.5, 1.5))))),
         "delta_lights": dict(scene=dict(delta_lights=(
             dict(kind="point", from_=(0.5, 1.5, -2.5), I=6.0),
             dict(kind="spot", from_=(-2.0, 2.0, -3.0), to=(0.0, 0.0, 0.0), I=(30.0, 24.0, 18.0), coneangle=25.0, conedelta=8.0),
             dict(kind="distant", from_=(1.0, 2.0, -3.0), to=(0.0, 0.0, 0.0), L=0.8)))),
         "delta_power": dict(scene=dict(delta_lights=(
             dict(kind="point", from_=(-1.0, 0.5, -2.8), I=9.0),
             dict(kind="spot", from_=(2.0, 2.5, -2.0), to=(0.2, -0.1, 0.0), I=40.0, coneangle=35.0, conedelta=35.0),
             dict(kind="distant", from_=(0.0, 1.0, -1.0), to=(0.0, 0.0, 0.0), L=(1.0, 0.9, 0.8))))),
         "filter_gaussian": dict(camera=dict(pixel_filter="gaussian")),
         "filter_mitchell": dict(camera=dict(pixel_filter="mitchell", max_sample_luminance=20.0)),
         "filter_sinc": dict(camera=dict(pixel_filter="sinc", sampler="halton")),
         "filter_aniso_crop": dict(camera=dict(pixel_filter="gaussian_aniso", crop_window=(0.21, 0.83, 0.1, 0.74))),
         "filter_triangle": dict(camera=dict(pixel_filter="triangle", lens_radius=0.05, focal_distance=4.5)),
         "filter_box1": dict(camera=dict(pixel_filter="box1")),
         "sphere_light": dict(scene=dict(spheres=(dict(center=(0.5, 2.5, -2.0), radius=0.3, emit=400.0),
                                                  dict(center=(-0.6, 0.2, -2.0), radius=0.5, material="plastic"))),
                              camera=dict(sampler="halton")),
         "sphere_power": dict(scene=dict(spheres=(dict(center=(0.0, 0.0, -2.4), radius=0.25, emit=150.0, two_sided=True),
                                                  dict(center=(2.2, -1.0, 0.0), radius=0.6, emit=30.0,
                                                       reverse_orientation=True),
                                                  dict(center=(-0.7, 0.5, -1.9), radius=0.35, material="glass_rough"))))}


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
