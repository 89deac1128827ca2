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
