"""Replays a scene dumped by pbrt_b200 (B200PT_DUMP_SCENE=<file>) through the ctypes harness: the exact descriptors
the C++ host hands to the C ABI, for the oracle (CPU) and for libb200pt (GPU).  Debugging aid, not a test.

  python tests/replay_dump.py <dump> oracle <out.pfm>     # CPU restatement
  python tests/replay_dump.py <dump> gpu <out.pfm>        # on a B200
"""
import ctypes as C
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


class DumpScene:
    def __init__(self, path, abi, scenes):
        raw = open(path, "rb").read()
        hdr = struct.unpack_from("<8q", raw, 0)
        assert hdr[0] == 0x3154504D55443042
        nt, nm, nl, ns, has_n, has_uv, stype = hdr[1:]
        off = 64

        def take(dtype, count):
            nonlocal off
            a = np.frombuffer(raw, dtype, count, off).copy()
            off += a.nbytes
            return a
        self.vertices = take("<f4", nt * 9).reshape(nt, 3, 3)
        self.material_id = take("<i4", nt)
        self.light_id = take("<i4", nt)
        self.flip = take("u1", nt)
        self.vertex_flags = take("u1", nt)
        self.normals = take("<f4", nt * 9) if has_n else None
        self.uvs = take("<f4", nt * 6) if has_uv else None
        self._materials = (abi.Material * nm).from_buffer_copy(raw, off)
        off += C.sizeof(abi.Material) * nm
        self._lights = (abi.AreaLight * max(nl, 1))()
        C.memmove(self._lights, raw[off:off + C.sizeof(abi.AreaLight) * nl], C.sizeof(abi.AreaLight) * nl)
        off += C.sizeof(abi.AreaLight) * nl
        self._spheres = (abi.Sphere * max(ns, 1))()
        C.memmove(self._spheres, raw[off:off + C.sizeof(abi.Sphere) * ns], C.sizeof(abi.Sphere) * ns)
        off += C.sizeof(abi.Sphere) * ns
        self.camera = abi.CameraDesc.from_buffer_copy(raw, off)
        off += C.sizeof(abi.CameraDesc)
        self.film = abi.FilmDesc.from_buffer_copy(raw, off)
        off += C.sizeof(abi.FilmDesc)
        self.integrator = abi.IntegratorDesc.from_buffer_copy(raw, off)
        self.integrator.n_bounded_media = 0  # the dump carries the host's pointers, not the tables behind them
        self.integrator.bounded_media = None
        self.integrator.sphere_medium = None
        off += C.sizeof(abi.IntegratorDesc)
        sm = struct.unpack_from("<6i", raw, off)
        self.n_triangles, self.n_lights, self.n_spheres, self.n_materials = nt, nl, ns, nm
        self.sampler_type, self.spp, self.sample_bounds = stype, sm[0], list(sm[1:5])
        self.abi, self.scenes = abi, scenes

    def desc(self):
        abi = self.abi
        d = abi.SceneDesc()
        d.n_triangles = self.n_triangles
        d.vertices = abi.ptr(self.vertices)
        d.material_id = abi.ptr(self.material_id)
        d.light_id = abi.ptr(self.light_id)
        d.flip_normal = abi.ptr(self.flip)
        d.n_materials = self.n_materials
        d.materials = C.cast(self._materials, C.POINTER(abi.Material))
        d.n_lights = self.n_lights
        d.lights = C.cast(self._lights, C.POINTER(abi.AreaLight))
        d.normals = abi.ptr(self.normals)
        d.uvs = abi.ptr(self.uvs)
        d.vertex_flags = abi.ptr(self.vertex_flags)
        d.n_spheres = self.n_spheres
        d.spheres = C.cast(self._spheres, C.POINTER(abi.Sphere))
        return d

    def setup(self):
        """A RenderSetup whose descriptors are the dumped ones (sampler tables from the golden fixtures)."""
        scenes, abi = self.scenes, self.abi
        res = self.film.full_resolution
        st = scenes.RenderSetup(res[0], res[1], self.spp, camera=self.camera,
                                sampler="halton" if self.sampler_type == abi.SAMPLER_HALTON else "sobol")
        st.film = self.film
        st.crop = list(self.film.cropped_bounds)
        st.sampler.sample_bounds[:] = self.sample_bounds
        st.sampler.samples_per_pixel = self.spp
        if self.sampler_type != abi.SAMPLER_HALTON:
            st._sobol_tables(self.sample_bounds)
        st.integrator = self.integrator
        return st


def main():
    pkg = g.load_package()
    from pbrt_v3_distributed_b200 import abi, scenes
    dump = DumpScene(sys.argv[1], abi, scenes)
    setup = dump.setup()
    print("dump: %d triangles, %d spheres, %d lights, %d materials, spp %d" %
          (dump.n_triangles, dump.n_spheres, dump.n_lights, dump.n_materials, dump.spp))
    if sys.argv[2] == "oracle":
        ob = g.load_oracle()
        o = ob.Oracle(abi, dump)
        film, _ = o.render(setup)
        rgb = o.film_rgb(setup, film)
    else:
        ctx = pkg.Context(0)
        scene = pkg.Scene(ctx, dump.desc(), keepalive=dump)
        r = pkg.Render(scene, setup)
        r.render_tiles()
        rgb = r.read_rgb()
    scenes.write_pfm(sys.argv[3], rgb)


if __name__ == "__main__":
    main()
