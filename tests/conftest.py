import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def abi(pkg):
    from pbrt_v3_distributed_b200 import abi as a
    return a


@pytest.fixture(scope="session")
def scenes(pkg):
    from pbrt_v3_distributed_b200 import scenes as s
    return s


@pytest.fixture(scope="session")
def ob():
    return graft.load_oracle()


@pytest.fixture(scope="session")
def probe_json():
    return json.load(open(os.path.join(GOLDEN, "probe.json")))


def hexf(v):
    return np.array([float.fromhex(x) for x in v], dtype=np.float64).astype(np.float32)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def golden_camera(abi, probe_json, w, h):
    rec = probe_json["cameras"]["%dx%d" % (w, h)]
    cam = abi.CameraDesc()
    cam.raster_to_camera[:] = list(hexf(rec["raster_to_camera"]))
    cam.camera_to_world[:] = list(hexf(rec["camera_to_world"]))
    cam.lens_radius = 0.0
    cam.focal_distance = 1e6
    cam.shutter_open = 0.0
    cam.shutter_close = 1.0
    return cam


HOSTCHECK_DIR = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="session")
def hostcheck(pkg):
    """The package's ctypes wrapper bound to tests/emu/libb200pt_hostcheck.so: the product's own api.cu / kernels.cu
    compiled as plain C++ (tests/emu/cuda_runtime.h).  TEST INFRASTRUCTURE -- a CPU pre-flight of the sources, never
    loaded by the package itself."""
    import subprocess
    import types
    lib = os.path.join(HOSTCHECK_DIR, "libb200pt_hostcheck.so")
    if os.environ.get("B200PT_HOSTCHECK_LIB"):
        # a differently compiled check build (ASan / UBSan run, profiles/r02/hostcheck_sanitizers.txt)
        lib = os.environ["B200PT_HOSTCHECK_LIB"]
    else:
        r = subprocess.run(["make", "-j", "8"], cwd=HOSTCHECK_DIR, capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(lib):
            pytest.fail("tests/emu does not build:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    init = os.path.join(graft.PKG_DIR, "__init__.py")
    src = open(init).read()
    marker = 'LIB_PATH = os.path.join(_HERE, "libb200pt.so")'
    assert marker in src
    mod = types.ModuleType("pbrt_v3_distributed_b200_hostcheck")
    mod.__file__ = init
    mod.__package__ = graft.PKG_NAME
    exec(compile(src.replace(marker, "LIB_PATH = %r" % lib), init, "exec"), mod.__dict__)
    return mod
