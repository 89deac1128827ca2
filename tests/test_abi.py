"""CPU tests: the C-ABI library loads, exports every symbol include/b200pt.h
declares, struct layouts match, and compute entry points fail loudly without a
GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "b200pt.h")).read()
    declared = set(re.findall(r"\b(b200pt_[a-z_0-9]+)\s*\(", header))
    declared -= {"b200pt_status"}
    assert declared, "no declarations parsed"
    nm = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (b200pt_[a-z_0-9]+)", nm))
    assert declared <= exported, "missing: %s" % sorted(declared - exported)
    assert not pkg.MISSING_SYMBOLS
    assert set(pkg.EXPORTED_SYMBOLS) == declared
    assert pkg.lib.b200pt_abi_version() == 6


def test_struct_sizes_match_header(abi):
    # sizes implied by include/b200pt.h on LP64
    assert C.sizeof(abi.Material) == 4 + 15 * 4 + 12 + 4
    assert C.sizeof(abi.AreaLight) == 116
    assert C.sizeof(abi.Sphere) == 188
    assert C.sizeof(abi.CameraDesc) == 144
    assert C.sizeof(abi.FilmDesc) == 48
    assert C.sizeof(abi.SamplerDesc) == 64
    assert C.sizeof(abi.IntegratorDesc) == 96 and C.sizeof(abi.Medium) == 40
    assert C.sizeof(abi.SceneDesc) == 168
    assert C.sizeof(abi.Instance) == 176
    assert abi.RAY_DTYPE.itemsize == 32 and abi.HIT_DTYPE.itemsize == 16


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.B200ptError) as e:
        pkg.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_host_preflight_traversal_matches_oracle(tmp_path):
    """The product's BVH build + traversal arithmetic, compiled for the host, against the oracle."""
    exe = str(tmp_path / "host_preflight")
    src = [os.path.join(ROOT, "tests", "host_preflight.cpp"),
           os.path.join(ROOT, "pbrt-v3-distributed_b200", "csrc", "wbvh_build.cpp")]
    r = subprocess.run(["g++", "-O2", "-std=gnu++17", "-ffp-contract=off", "-pthread", *src, "-o", exe,
                        "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                        "-Wl,-rpath," + os.path.join(ROOT, "oracle")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    for args in (["20000", "30000", "1"], ["1", "500", "2"], ["50", "5000", "3"]):
        r = subprocess.run([exe] + args, capture_output=True, text=True)
        assert r.returncode == 0 and "PREFLIGHT OK" in r.stdout, r.stdout + r.stderr


def test_device_sincos_matches_host_libm(tmp_path, abi, ob):
    """pt_sincos.cuh (host-compiled) == the libm sinf/cosf the reference calls, on the hot-path range."""
    src = tmp_path / "sc.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cmath>
#include "pt_sincos.cuh"
int main() { long bad = 0, n = 0;
  for (float x = -3.2f; x < 3.2f; x = std::nextafter(x, 4.f)) { if ((++n & 63) != 0) continue;
    if (b200pt::float_as_uint(sinf(x)) != b200pt::float_as_uint(b200pt::pt_sinf(x))) ++bad;
    if (b200pt::float_as_uint(cosf(x)) != b200pt::float_as_uint(b200pt::pt_cosf(x))) ++bad; }
  printf("%ld %ld\n", n, bad); return bad != 0; }
''')
    exe = str(tmp_path / "sc")
    inc = os.path.join(ROOT, "pbrt-v3-distributed_b200", "csrc")
    r = subprocess.run(["g++", "-O2", "-std=gnu++17", "-ffp-contract=off", "-I" + inc, str(src), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
