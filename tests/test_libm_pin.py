"""CPU test: the device's expf / logf restatements (pbrt-v3-distributed_b200/csrc/pt_explog.cuh, groundwork for media:
free-flight sampling and transmittance must round like the host libm the reference calls) against std::exp / std::log.
tests/libm_pin.cpp checks EVERY float bit pattern (about 15 s on 8 cores; 0 mismatches, DESIGN.md "Numerics"); this test
runs every 13th pattern to stay quick."""
import os
import subprocess

from conftest import ROOT


def test_expf_logf_round_like_the_host_libm(tmp_path):
    exe = str(tmp_path / "libm_pin")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", os.path.join(ROOT, "tests", "libm_pin.cpp"),
                    "-o", exe], check=True)
    r = subprocess.run([exe, "13"], capture_output=True, text=True)
    assert r.returncode == 0 and "expf: 0 mismatches" in r.stdout and "logf: 0 mismatches" in r.stdout, r.stdout
