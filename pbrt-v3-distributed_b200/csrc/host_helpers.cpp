// host_helpers.cpp -- host-side math a caller outside pbrt needs to fill the
// C-ABI descriptors exactly like the reference would (no device work).
//
// b200pt_host_perspective_camera follows what pbrtLookAt + pbrtCamera +
// ProjectiveCamera do to arrive at RasterToCamera / CameraToWorld:
//   api.cpp:993-1001 (curTransform = curTransform * LookAt), api.cpp:1131-1146
//   (CameraToWorld = Inverse(curTransform)), transform.cpp:203-249 (LookAt),
//   transform.cpp:303-318 (Perspective(fov, 1e-2, 1000)), camera.h:98-111
//   (ScreenToRaster, RasterToScreen, RasterToCamera), perspective.cpp:227-273
//   (screen window from the frame aspect ratio), transform.cpp:82-139 (Inverse).
// All arithmetic is float in the reference's order so the matrices are
// bit-identical (tests/test_host_helpers.py checks against matrices dumped
// from the reference).
#include "pt_sphere.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>

#include "../../include/b200pt.h"
#include "b200pt_internal.h"

namespace {

struct M4 {
    float m[4][4];
};

M4 Identity() {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = i == j ? 1.f : 0.f;
    return r;
}

// Matrix4x4::Mul, transform.h:88-96
M4 Mul(const M4 &a, const M4 &b) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] +
                        a.m[i][3] * b.m[3][j];
    return r;
}

// Gauss-Jordan with full pivoting, transform.cpp:82-139
bool Inverse(const M4 &m, M4 *out) {
    int indxc[4], indxr[4];
    int ipiv[4] = {0, 0, 0, 0};
    float minv[4][4];
    memcpy(minv, m.m, sizeof(minv));
    for (int i = 0; i < 4; i++) {
        int irow = 0, icol = 0;
        float big = 0.f;
        for (int j = 0; j < 4; j++) {
            if (ipiv[j] != 1) {
                for (int k = 0; k < 4; k++) {
                    if (ipiv[k] == 0) {
                        if (std::abs(minv[j][k]) >= big) {
                            big = float(std::abs(minv[j][k]));
                            irow = j;
                            icol = k;
                        }
                    } else if (ipiv[k] > 1)
                        return false;
                }
            }
        }
        ++ipiv[icol];
        if (irow != icol)
            for (int k = 0; k < 4; ++k) std::swap(minv[irow][k], minv[icol][k]);
        indxr[i] = irow;
        indxc[i] = icol;
        if (minv[icol][icol] == 0.f) return false;
        float pivinv = 1. / minv[icol][icol];
        minv[icol][icol] = 1.;
        for (int j = 0; j < 4; j++) minv[icol][j] *= pivinv;
        for (int j = 0; j < 4; j++) {
            if (j != icol) {
                float save = minv[j][icol];
                minv[j][icol] = 0;
                for (int k = 0; k < 4; k++) minv[j][k] -= minv[icol][k] * save;
            }
        }
    }
    for (int j = 3; j >= 0; j--) {
        if (indxr[j] != indxc[j])
            for (int k = 0; k < 4; k++) std::swap(minv[k][indxr[j]], minv[k][indxc[j]]);
    }
    memcpy(out->m, minv, sizeof(minv));
    return true;
}

struct V {
    float x, y, z;
};
float Len(V v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }
V Norm(V v) {
    float inv = (float)1 / Len(v);
    return V{v.x * inv, v.y * inv, v.z * inv};
}
V CrossD(V a, V b) {  // geometry.h:957-963, evaluated in double
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return V{(float)((ay * bz) - (az * by)), (float)((az * bx) - (ax * bz)), (float)((ax * by) - (ay * bx))};
}

// Transform = (m, mInv) pair, transform.h:115-130,251-253
struct Xf {
    M4 m, mInv;
};
Xf operator*(const Xf &a, const Xf &b) { return Xf{Mul(a.m, b.m), Mul(b.mInv, a.mInv)}; }
Xf InverseXf(const Xf &t) { return Xf{t.mInv, t.m}; }
// transform.cpp:149-153
Xf ScaleXf(float x, float y, float z) {
    Xf r{Identity(), Identity()};
    r.m.m[0][0] = x;
    r.m.m[1][1] = y;
    r.m.m[2][2] = z;
    r.mInv.m[0][0] = 1 / x;
    r.mInv.m[1][1] = 1 / y;
    r.mInv.m[2][2] = 1 / z;
    return r;
}
// transform.cpp:141-147
Xf TranslateXf(float dx, float dy, float dz) {
    Xf r{Identity(), Identity()};
    r.m.m[0][3] = dx;
    r.m.m[1][3] = dy;
    r.m.m[2][3] = dz;
    r.mInv.m[0][3] = -dx;
    r.mInv.m[1][3] = -dy;
    r.mInv.m[2][3] = -dz;
    return r;
}

}  // namespace

extern "C" int b200pt_host_perspective_camera(const float eye[3], const float look[3], const float up[3],
                                              float fov, int32_t xres, int32_t yres,
                                              b200pt_camera_desc *out) {
    if (!eye || !look || !up || !out || xres <= 0 || yres <= 0)
        return b200pt_fail(B200PT_ERR_INVALID, "host_perspective_camera: bad argument");
    // LookAt, transform.cpp:203-237: returns Transform(Inverse(cameraToWorld), cameraToWorld)
    V pos{eye[0], eye[1], eye[2]};
    V dir = Norm(V{look[0] - eye[0], look[1] - eye[1], look[2] - eye[2]});
    V upn = Norm(V{up[0], up[1], up[2]});
    if (Len(CrossD(upn, dir)) == 0)
        return b200pt_fail(B200PT_ERR_INVALID, "host_perspective_camera: up parallel to view");
    V right = Norm(CrossD(upn, dir));
    V newUp = CrossD(dir, right);
    M4 c2w;
    memset(&c2w, 0, sizeof(c2w));
    c2w.m[0][3] = pos.x;
    c2w.m[1][3] = pos.y;
    c2w.m[2][3] = pos.z;
    c2w.m[3][3] = 1;
    c2w.m[0][0] = right.x;
    c2w.m[1][0] = right.y;
    c2w.m[2][0] = right.z;
    c2w.m[0][1] = newUp.x;
    c2w.m[1][1] = newUp.y;
    c2w.m[2][1] = newUp.z;
    c2w.m[0][2] = dir.x;
    c2w.m[1][2] = dir.y;
    c2w.m[2][2] = dir.z;
    Xf lookAt;
    lookAt.mInv = c2w;
    if (!Inverse(c2w, &lookAt.m))
        return b200pt_fail(B200PT_ERR_INVALID, "host_perspective_camera: singular LookAt");
    // api.cpp:993-1001: curTransform = curTransform * lookAt (curTransform starts as identity);
    // api.cpp:1131-1146: CameraToWorld = Inverse(curTransform)
    Xf ctm = Xf{Identity(), Identity()} * lookAt;
    Xf cameraToWorld = InverseXf(ctm);
    // Perspective(fov, 1e-2f, 1000.f), transform.cpp:303-312
    float n = 1e-2f, f = 1000.f;
    Xf persp;
    persp.m = Identity();
    persp.m.m[2][2] = f / (f - n);
    persp.m.m[2][3] = -f * n / (f - n);
    persp.m.m[3][2] = 1;
    persp.m.m[3][3] = 0;
    if (!Inverse(persp.m, &persp.mInv))
        return b200pt_fail(B200PT_ERR_INVALID, "host_perspective_camera: singular projection");
    const float PiF = 3.14159265358979323846;
    float invTanAng = 1 / std::tan(((PiF / 180) * fov) / 2);  // pbrt.h:318 Radians
    Xf cameraToScreen = ScaleXf(invTanAng, invTanAng, 1) * persp;
    // screen window, perspective.cpp:238-251
    float frame = float(xres) / float(yres);
    float sx0, sx1, sy0, sy1;
    if (frame > 1.f) {
        sx0 = -frame;
        sx1 = frame;
        sy0 = -1.f;
        sy1 = 1.f;
    } else {
        sx0 = -1.f;
        sx1 = 1.f;
        sy0 = -1.f / frame;
        sy1 = 1.f / frame;
    }
    // camera.h:99-106
    Xf screenToRaster = ScaleXf((float)xres, (float)yres, 1) *
                        ScaleXf(1 / (sx1 - sx0), 1 / (sy0 - sy1), 1) * TranslateXf(-sx0, -sy1, 0);
    Xf rasterToScreen = InverseXf(screenToRaster);
    Xf rasterToCamera = InverseXf(cameraToScreen) * rasterToScreen;
    memcpy(out->raster_to_camera, rasterToCamera.m.m, sizeof(float) * 16);
    memcpy(out->camera_to_world, cameraToWorld.m.m, sizeof(float) * 16);
    out->lens_radius = 0.f;
    out->focal_distance = 1e6f;
    out->shutter_open = 0.f;
    out->shutter_close = 1.f;
    return B200PT_OK;
}

// microfacet.h:123-128
extern "C" float b200pt_host_roughness_to_alpha(float roughness) {
    roughness = std::max(roughness, (float)1e-3);
    float x = std::log(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x +
           0.000640711f * x * x * x * x;
}

// OrenNayar::OrenNayar, reflection.h:414-420, after matte.cpp:54's Clamp(sigma, 0, 90)
extern "C" void b200pt_host_oren_nayar(float sigma, float *A, float *B) {
    sigma = sigma < 0 ? 0.f : (sigma > 90 ? 90.f : sigma);
    const float PiF = 3.14159265358979323846;
    sigma = (PiF / 180) * sigma;  // Radians
    float sigma2 = sigma * sigma;
    *A = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
    *B = 0.45f * sigma2 / (sigma2 + 0.09f);
}

// Sphere::Sphere, sphere.h:49-61 (std::acos through the restated libm routine of pt_sphere.cuh)
extern "C" void b200pt_host_sphere_params(float radius, float zmin, float zmax, float phimax_degrees, float out[5]) {
    using namespace b200pt;
    out[0] = pt_clamp(pt_min(zmin, zmax), -radius, radius);
    out[1] = pt_clamp(pt_max(zmin, zmax), -radius, radius);
    out[2] = pt_acosf(pt_clamp(pt_min(zmin, zmax) / radius, -1.f, 1.f));
    out[3] = pt_acosf(pt_clamp(pt_max(zmin, zmax) / radius, -1.f, 1.f));
    out[4] = (PT_PI / 180) * pt_clamp(phimax_degrees, 0.f, 360.f);
}

// CreateSpotLight (lights/spot.cpp:104-124) at the identity CTM + SpotLight ctor (:42-51).
// light2world = Translate(from) * Inverse(dirToZ), so WorldToLight.m = dirToZ.m * Translate(-from).m
// (Transform::operator*, transform.cpp:222-225: no Gauss-Jordan inverse is involved on this side) and
// pLight = LightToWorld(0,0,0) = from.
extern "C" void b200pt_host_spot_light(const float from[3], const float to[3], float coneangle, float conedelta,
                                       b200pt_area_light *out) {
    using namespace b200pt;
    const V3 dir = normalize(mk(to[0] - from[0], to[1] - from[1], to[2] - from[2]));
    V3 du, dv;
    coordinate_system(dir, &du, &dv);
    const float dz[16] = {du.x, du.y, du.z, 0.f, dv.x, dv.y, dv.z, 0.f, dir.x, dir.y, dir.z, 0.f, 0.f, 0.f, 0.f, 1.f};
    const float ti[16] = {1.f, 0.f, 0.f, -from[0], 0.f, 1.f, 0.f, -from[1], 0.f, 0.f, 1.f, -from[2], 0.f, 0.f, 0.f, 1.f};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)  // Matrix4x4::Mul, transform.cpp:98-106
            out->world_to_light[4 * i + j] =
                dz[4 * i] * ti[j] + dz[4 * i + 1] * ti[4 + j] + dz[4 * i + 2] * ti[8 + j] + dz[4 * i + 3] * ti[12 + j];
    out->kind = B200PT_LIGHT_SPOT;
    out->position[0] = from[0];
    out->position[1] = from[1];
    out->position[2] = from[2];
    const float totalWidth = coneangle, falloffStart = coneangle - conedelta;
    out->cos_total_width = pt_cosf((PT_PI / 180) * totalWidth);
    out->cos_falloff_start = pt_cosf((PT_PI / 180) * falloffStart);
}
