"""Regenerates tests/golden/* from the UNMODIFIED reference (oracle/_ref/).

Run here (the container that has /root/reference):  python tests/golden/make_golden.py
Fixtures:
  sobol_tables.bin   generator matrices dumped from core/sobolmatrices.cpp (first 256 dims)
  probe.json         camera matrices, Sobol' sample streams, camera rays, material constants
                     (hex floats) from PerspectiveCamera / SobolSampler / MetalMaterial
  isect_*.npy        rays + BVHAccel::Intersect / IntersectP answers on a 3000-triangle soup
  render_*.pfm       pbrt_ref renders of small soup scenes (all four materials, uniform/power)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
ob = g.load_oracle()
from pbrt_v3_distributed_b200 import abi, scenes  # noqa: E402

CAMERAS = [(64, 64), (96, 64), (48, 80), (1024, 1024), (1920, 1080), (3840, 2160)]
RENDERS = {
    # name: (n_tris, materials, xres, yres, spp, maxdepth, strategy, n_lights)
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


HALTON_CASES = [((0, 0, 700, 700), 8, 345, 678, 5, 0, 100), ((0, 0, 64, 48), 16, 3, 47, 15, 0, 120),
                ((0, 0, 1920, 1080), 1024, 1919, 1079, 1023, 0, 128), ((13, 7, 90, 50), 5, 13, 7, 0, 0, 64),
                ((0, 0, 100, 100), 3, 99, 0, 2, 0, 64), ((0, 0, 1, 1), 4, 0, 0, 3, 0, 16)]


def halton_fixtures():
    """HaltonSampler: the digit permutations of the first 128 prime bases and sample streams."""
    ob.probe("haltonperms", os.path.join(HERE, "halton_perms.bin"), 128)
    out = []
    for (bounds, spp, px, py, sample, dim0, n) in HALTON_CASES:
        vals = ob.probe("halton", *bounds, spp, px, py, sample, dim0, n).split()
        out.append({"bounds": bounds, "spp": spp, "px": px, "py": py, "sample": sample, "dim0": dim0,
                    "index": int(vals[0]), "values": vals[1:]})
    json.dump(out, open(os.path.join(HERE, "probe_halton.json"), "w"), indent=1)


FILTERS = {  # key: (probe args, the PixelFilter line of the .pbrt twin)
    "gaussian": (("gaussian", 2, 2, 2), 'PixelFilter "gaussian" "float xwidth" [2] "float ywidth" [2] "float alpha" [2]'),
    "gaussian_aniso": (("gaussian", 1.5, 2.5, 1), 'PixelFilter "gaussian" "float xwidth" [1.5] "float ywidth" [2.5] "float alpha" [1]'),
    "mitchell": (("mitchell", 2, 2), 'PixelFilter "mitchell" "float xwidth" [2] "float ywidth" [2]'),
    "triangle": (("triangle", 2, 2), 'PixelFilter "triangle" "float xwidth" [2] "float ywidth" [2]'),
    "sinc": (("sinc", 4, 4, 3), 'PixelFilter "sinc" "float xwidth" [4] "float ywidth" [4] "float tau" [3]'),
    "box1": (("box", 1, 1), 'PixelFilter "box" "float xwidth" [1] "float ywidth" [1]'),
}


def filter_fixtures():
    out = {}
    for key, (args, line) in FILTERS.items():
        txt = ob.probe("filtertable", *args).splitlines()
        out[key] = {"pbrt": line, "radius": txt[0].split()[1:], "table": txt[1:257]}
    json.dump(out, open(os.path.join(HERE, "filter_tables.json"), "w"), indent=0)


def main():
    ob.probe("tables", os.path.join(HERE, "sobol_tables.bin"), 256)
    filter_fixtures()
    halton_fixtures()
    out = {"cameras": {}, "sobol": [], "camrays": []}
    cam_args = [0, 0, -4.5, 0, 0, 0, 0, 1, 0, 35]
    for (w, h) in CAMERAS:
        txt = ob.probe("camera", *cam_args, w, h).splitlines()
        rec = {}
        for line in txt:
            k, *v = line.split()
            rec[k] = v
        out["cameras"]["%dx%d" % (w, h)] = rec
    for (bounds, spp, px, py, sample, dim0, n) in [((0, 0, 64, 64), 16, 3, 5, 2, 0, 48),
                                                   ((0, 0, 1024, 1024), 256, 1023, 517, 255, 0, 48),
                                                   ((0, 0, 1920, 1080), 1024, 1919, 1079, 1023, 0, 140),
                                                   ((0, 0, 1920, 1080), 1024, 0, 0, 0, 0, 16),
                                                   ((0, 0, 3840, 2160), 4096, 2000, 1000, 4095, 100, 40)]:
        vals = ob.probe("sobol", *bounds, spp, px, py, sample, dim0, n).split()
        out["sobol"].append({"bounds": bounds, "spp": spp, "px": px, "py": py, "sample": sample, "dim0": dim0,
                             "values": vals})
    for (w, h, spp, px, py, n) in [(64, 64, 16, 3, 5, 16), (1920, 1080, 1024, 1900, 1000, 8)]:
        lines = ob.probe("camrays", *cam_args, w, h, spp, px, py, n).splitlines()
        out["camrays"].append({"res": [w, h], "spp": spp, "px": px, "py": py, "rays": [l.split() for l in lines]})
    consts = {}
    for line in ob.probe("consts").splitlines():
        k, *v = line.split()
        consts.setdefault(k, []).append(v)
    out["consts"] = consts
    json.dump(out, open(os.path.join(HERE, "probe.json"), "w"), indent=1)

    # ray/triangle + BVH answers from the real BVHAccel
    arr = scenes.SceneArrays(3000, materials=("matte",), soup_version=1, seed=99)
    rng = np.random.default_rng(5)
    n = 6000
    rays = np.zeros(n, dtype=abi.RAY_DTYPE)
    rays["o"] = rng.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[: n // 2] /= np.linalg.norm(d[: n // 2], axis=1, keepdims=True)  # half unit, half unnormalised
    rays["d"] = d
    rays["t_max"] = np.inf
    rays["t_max"][n // 3: n // 2] = rng.uniform(0.05, 1.0, n // 2 - n // 3).astype(np.float32)
    # axis-aligned directions exercise the zero-component paths of the slab and watertight tests
    rays["d"][:30] = np.tile(np.eye(3, dtype=np.float32), (10, 1))
    tmp = "/tmp/golden_isect"
    os.makedirs(tmp, exist_ok=True)
    arr.vertices.tofile(os.path.join(tmp, "tris.f32"))
    rays.tofile(os.path.join(tmp, "rays.bin"))
    ob.probe("intersect", os.path.join(tmp, "tris.f32"), os.path.join(tmp, "rays.bin"), os.path.join(tmp, "out.bin"))
    res = np.fromfile(os.path.join(tmp, "out.bin"),
                      dtype=[("tri", "<i4"), ("t", "<f4"), ("p", "<f4", 3), ("n", "<f4", 3), ("perr", "<f4", 3),
                             ("occluded", "<i4")])
    np.save(os.path.join(HERE, "isect_rays.npy"), rays)
    np.save(os.path.join(HERE, "isect_ref.npy"), res)

    for name, (nt, mats, w, h, spp, depth, strat, nl) in RENDERS.items():
        ex = EXTRA.get(name, {})
        arr = scenes.SceneArrays(nt, materials=mats, soup_version=1, n_lights=nl, **ex.get("scene", {}))
        path = scenes.write_pbrt("/tmp/golden_render", "render_" + name, arr, w, h, spp, max_depth=depth,
                                 strategy=strat, **ex.get("camera", {}))
        ob.run_pbrt_ref(path, threads=1 if name.startswith("filter_") else None)
        os.replace(os.path.join("/tmp/golden_render", "render_%s.pfm" % name),
                   os.path.join(HERE, "render_%s.pfm" % name))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
