"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
PBRT_REF = os.path.join(REF_DIR, "pbrt_ref")
REF_PROBE = os.path.join(REF_DIR, "ref_probe")


def have_reference():
    return os.path.exists(PBRT_REF) and os.path.exists(REF_PROBE)


LIB_SPECTRAL_PATH = os.path.join(_HERE, "liboracle_spectral.so")
PBRT_REF_SPECTRAL = os.path.join(REF_DIR, "pbrt_ref_spectral")
REF_PROBE_SPECTRAL = os.path.join(REF_DIR, "ref_probe_spectral")


def load(abi, spectral=False):
    """abi = the product package's ctypes mirror of include/b200pt.h.  spectral: the 60-bin SampledSpectrum build."""
    lib = C.CDLL(LIB_SPECTRAL_PATH if spectral else LIB_PATH)
    lib.oracle_spectral_register.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.oracle_spectral_set_cie.argtypes = [C.POINTER(C.c_float)] * 3
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.oracle_scene_create.restype = vp
    lib.oracle_scene_create.argtypes = [C.POINTER(abi.SceneDesc)]
    lib.oracle_scene_destroy.argtypes = [vp]
    for n in ("oracle_trace_closest", "oracle_trace_any", "oracle_trace_closest_brute"):
        getattr(lib, n).argtypes = [vp, vp, vp, i64]
    lib.oracle_render.argtypes = [vp, C.POINTER(abi.CameraDesc), C.POINTER(abi.FilmDesc),
                                  C.POINTER(abi.SamplerDesc), C.POINTER(abi.IntegratorDesc), vp, i64,
                                  C.c_int, vp, C.POINTER(abi.Stats)]
    lib.oracle_film_rgb.argtypes = [C.POINTER(abi.FilmDesc), vp, vp]
    lib.oracle_sobol.argtypes = [C.POINTER(abi.SamplerDesc), i32, i32, i64, i32, i32, vp]
    lib.oracle_camera_rays.argtypes = [C.POINTER(abi.CameraDesc), C.POINTER(abi.SamplerDesc), i32, i32, i32, vp]
    lib.oracle_pixel_samples.argtypes = [vp, C.POINTER(abi.CameraDesc), C.POINTER(abi.FilmDesc),
                                         C.POINTER(abi.SamplerDesc), C.POINTER(abi.IntegratorDesc), i32, i32, vp]
    lib.oracle_libm_sinf.restype = C.c_float
    lib.oracle_libm_sinf.argtypes = [C.c_float]
    lib.oracle_libm_cosf.restype = C.c_float
    lib.oracle_libm_cosf.argtypes = [C.c_float]
    lib.oracle_set_volpath.restype = None
    lib.oracle_set_volpath.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float]
    return lib


def set_medium_boundaries(lib, scene_arrays):
    """Groundwork: spheres of the scene with a `boundary` spec are null-material surfaces around a homogeneous medium."""
    specs = [(k, sp["boundary"]) for k, sp in enumerate(getattr(scene_arrays, "sphere_specs", ())) if sp.get("boundary")]
    n = len(specs)
    idx = (C.c_int * max(n, 1))(*[k for k, _ in specs])
    sa = (C.c_float * max(3 * n, 1))(*[v for _, b_ in specs for v in b_["sigma_a"]])
    ss = (C.c_float * max(3 * n, 1))(*[v for _, b_ in specs for v in b_["sigma_s"]])
    g = (C.c_float * max(n, 1))(*[b_.get("g", 0.0) for _, b_ in specs])
    lib.oracle_set_medium_boundaries.restype = None
    lib.oracle_set_medium_boundaries.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_set_medium_boundaries(n, idx, sa, ss, g)


def set_volpath(lib, enabled, medium=None):
    """VolPathIntegrator instead of PathIntegrator for the renders that follow (process-wide switch of the oracle library);
    medium = dict(sigma_a=(r, g, b), sigma_s=(r, g, b), g=...) is a homogeneous medium around the whole scene."""
    f3 = C.c_float * 3
    m = medium or {}
    lib.oracle_set_volpath(int(bool(enabled)), int(medium is not None), f3(*m.get("sigma_a", (0, 0, 0))),
                           f3(*m.get("sigma_s", (0, 0, 0))), float(m.get("g", 0.0)))


class Oracle:
    def __init__(self, abi, scene_arrays, spectral_tables=None):
        """spectral_tables: dict from tests/golden/spectral_tables.json -> the SampledSpectrum build of the oracle."""
        self.abi = abi
        described = getattr(scene_arrays, "cie_xyz", None) is not None  # SceneArrays.attach_spectral(): tables in the descriptor
        self.lib = load(abi, spectral=spectral_tables is not None or described)
        if spectral_tables is not None:
            f60 = C.c_float * 60

            def arr(v):
                return f60(*[float.fromhex(x) for x in v])
            self.lib.oracle_spectral_set_cie(arr(spectral_tables["cie_x"]), arr(spectral_tables["cie_y"]),
                                             arr(spectral_tables["cie_z"]))
            for rgb, spec in spectral_tables["spectra"]:
                self.lib.oracle_spectral_register((C.c_float * 3)(*[float.fromhex(x) for x in rgb]), arr(spec))
        self.arrays = scene_arrays
        d = scene_arrays.desc()
        self.h = self.lib.oracle_scene_create(C.byref(d))

    def trace_closest(self, rays, brute=False):
        rays = np.ascontiguousarray(rays, dtype=self.abi.RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=self.abi.HIT_DTYPE)
        fn = self.lib.oracle_trace_closest_brute if brute else self.lib.oracle_trace_closest
        fn(self.h, self.abi.ptr(rays), self.abi.ptr(hits), len(rays))
        return hits

    def trace_any(self, rays):
        rays = np.ascontiguousarray(rays, dtype=self.abi.RAY_DTYPE)
        occ = np.zeros(len(rays), dtype=np.uint8)
        self.lib.oracle_trace_any(self.h, self.abi.ptr(rays), self.abi.ptr(occ), len(rays))
        return occ

    def render(self, setup, tiles=None, n_tiles=-1, threads=None):
        cb = setup.film.cropped_bounds
        w, h = cb[2] - cb[0], cb[3] - cb[1]
        film = np.zeros((h, w, 4), np.float32)
        stats = self.abi.Stats()
        t = None if tiles is None else np.ascontiguousarray(tiles, dtype=np.int32)
        self.lib.oracle_render(self.h, C.byref(setup.camera), C.byref(setup.film), C.byref(setup.sampler),
                               C.byref(setup.integrator), self.abi.ptr(t), len(t) if t is not None else n_tiles,
                               threads or os.cpu_count(), self.abi.ptr(film), C.byref(stats))
        return film, {k: getattr(stats, k) for k, _ in self.abi.Stats._fields_}

    def film_rgb(self, setup, film):
        rgb = np.zeros(film.shape[:2] + (3,), np.float32)
        self.lib.oracle_film_rgb(C.byref(setup.film), self.abi.ptr(film), self.abi.ptr(rgb))
        return rgb

    def sobol(self, setup, px, py, sample, dim0, n):
        out = np.zeros(n, np.float32)
        self.lib.oracle_sobol(C.byref(setup.sampler), px, py, sample, dim0, n, self.abi.ptr(out))
        return out

    def camera_rays(self, setup, px, py, n):
        out = np.zeros(n, dtype=self.abi.RAY_DTYPE)
        self.lib.oracle_camera_rays(C.byref(setup.camera), C.byref(setup.sampler), px, py, n, self.abi.ptr(out))
        return out

    def pixel_samples(self, setup, px, py):
        out = np.zeros((setup.sampler.samples_per_pixel, 3), np.float32)
        self.lib.oracle_pixel_samples(self.h, C.byref(setup.camera), C.byref(setup.film), C.byref(setup.sampler),
                                      C.byref(setup.integrator), px, py, self.abi.ptr(out))
        return out

    def close(self):
        if self.h:
            self.lib.oracle_scene_destroy(self.h)
            self.h = None


def run_pbrt_ref(pbrt_file, threads=None, quiet=True, timeout=3600, spectral=False):
    """Runs the unmodified reference CLI (spectral: its SampledSpectrum build); returns its stdout (stats + profile)."""
    cmd = [PBRT_REF_SPECTRAL if spectral else PBRT_REF, "--nthreads", str(threads or os.cpu_count()),
           os.path.basename(pbrt_file)]
    r = subprocess.run(cmd, cwd=os.path.dirname(pbrt_file), capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("pbrt_ref failed: " + r.stderr[-2000:])
    return r.stdout + r.stderr


def probe(*args, spectral=False):
    r = subprocess.run([REF_PROBE_SPECTRAL if spectral else REF_PROBE] + [str(a) for a in args], capture_output=True,
                       text=True)
    if r.returncode != 0:
        raise RuntimeError("ref_probe failed: " + r.stderr[-2000:])
    return r.stdout
