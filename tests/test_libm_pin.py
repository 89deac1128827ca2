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


def test_device_henyey_greenstein_passes_the_references_hg_tests(tmp_path):
    """src/tests/hg.cpp restated for the device's phase function (pt_core.cuh phase_hg / hg_sample_p, compiled for the host):
    SamplingMatch (the sampled pdf equals p(wo, wi) within 1e-4), SamplingOrientationForward / Backward (g = +-0.95),
    Normalized (the mean of p over uniform directions is 1/4pi within 1e-3)."""
    src = tmp_path / "hg.cpp"
    src.write_text(r'''
#include <cmath>
#include <cstdio>
#include <random>
#include "pt_core.cuh"
using namespace b200pt;
static std::mt19937 rng(7);
static float U() { return std::uniform_real_distribution<float>(0.f, 1.f)(rng); }
static V3 uniformSphere() {  // sampling.cpp:132-137
    float u0 = U(), u1 = U();
    float z = 1 - 2 * u0, r = std::sqrt(std::max(0.f, 1 - z * z)), phi = 2 * PT_PI * u1;
    return mk(r * std::cos(phi), r * std::sin(phi), z);
}
int main() {
    int fail = 0;
    for (float g = -.75f; g <= 0.75f; g += 0.25f)  // SamplingMatch
        for (int i = 0; i < 100; ++i) {
            V3 wo = uniformSphere(), wi;
            float u[2] = {U(), U()};
            float p0 = hg_sample_p(g, wo, &wi, u);
            if (!(std::fabs(p0 - phase_hg(dot(wo, wi), g)) <= 1e-4f)) { ++fail; printf("SamplingMatch g=%g: %g vs %g\n", g, p0, phase_hg(dot(wo, wi), g)); }
        }
    for (float g : {0.95f, -0.95f}) {  // SamplingOrientationForward / Backward
        int nForward = 0, nBackward = 0;
        for (int i = 0; i < 100; ++i) {
            V3 wi;
            float u[2] = {U(), U()};
            hg_sample_p(g, mk(-1.f, 0.f, 0.f), &wi, u);
            (wi.x > 0 ? nForward : nBackward)++;
        }
        if (g > 0 ? !(nForward >= 10 * nBackward) : !(nBackward >= 10 * nForward)) { ++fail; printf("orientation g=%g: %d / %d\n", g, nForward, nBackward); }
    }
    for (float g = -.75f; g <= 0.75f; g += 0.25f) {  // Normalized
        V3 wo = uniformSphere();
        double sum = 0;
        const int n = 100000;
        for (int i = 0; i < n; ++i) sum += phase_hg(dot(wo, uniformSphere()), g);
        if (!(std::fabs(sum / n - 1. / (4. * 3.14159265358979323846)) <= 1e-3)) { ++fail; printf("Normalized g=%g: %g\n", g, sum / n); }
    }
    printf(fail ? "HG FAILED\n" : "HG OK\n");
    return fail != 0;
}
''')
    exe = str(tmp_path / "hg")
    inc = os.path.join(ROOT, "pbrt-v3-distributed_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I" + inc, str(src), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "HG OK" in r.stdout, r.stdout
