"""CPU test of bench.py's reference arm (`--impl reference`): it must map none of the repo's libraries, print the contract's
JSON line and end within its wall-clock budget whatever --steps / --warmup ask for (each step is one launch of the
unmodified reference: scene load + BVH build + render, DESIGN.md "Measurement")."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

PBRT_REF = os.path.join(ROOT, "oracle", "_ref", "pbrt_ref")


@pytest.mark.skipif(not os.path.exists(PBRT_REF), reason="oracle/_ref/pbrt_ref is built where /root/reference exists")
def test_reference_arm_is_bounded_and_maps_no_repo_library():
    env = dict(os.environ, B200PT_REF_BUDGET_S="6")
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'small', '--steps', '20', '--warmup', '5'];"
            "runpy.run_path('bench.py', run_name='not_main'); import bench; rc = bench.main();"
            "maps = open('/proc/self/maps').read(); assert 'libb200pt' not in maps, 'the reference arm mapped libb200pt.so'; sys.exit(rc)")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["value"] > 0
    assert line["steps"] == 20 and line["warmup"] == 5
    # 25 launches do not fit a 6 s budget: fewer are run, at least one of them timed, and the line says so
    assert 1 <= line["steps_run"] < 20 and line["warmup_run"] >= 1
    assert line["cpu_baseline"]["kind"] == "reference" and "budget" in line["cpu_baseline"]["sample"]
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
