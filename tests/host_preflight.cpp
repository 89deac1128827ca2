// tests/host_preflight.cpp -- CPU pre-flight of the product's device arithmetic.
//
// Compiles the product's host+device headers (pt_core.cuh, wbvh_traverse.cuh)
// and the host BVH builder as plain C++ and checks them against the oracle
// (oracle/liboracle.so) BEFORE GPU time is spent:
//   * 7-wide BVH build + traversal vs the oracle's closest / any hit answers (bit-exact t)
//   * Sobol' sample stream and camera rays vs the oracle
//   * the on-device BVH builder's per-element steps (lbvh.cuh) run sequentially: structure check + traversal vs the oracle
//   * Sphere::Intersect / IntersectP arithmetic (EFloat quadratic, error-tracking transforms) vs the oracle,
//     and the restated acosf vs the host libm
// This binary is test infrastructure: it is never linked into libb200pt.so and
// the product has no CPU path.  Usage: host_preflight <n_tris> <n_rays> <seed>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../oracle/pt_oracle.h"
#include "../pbrt-v3-distributed_b200/csrc/wbvh.h"
#include "../pbrt-v3-distributed_b200/csrc/wbvh_traverse.cuh"
#include "../pbrt-v3-distributed_b200/csrc/lbvh.cuh"
#include "../pbrt-v3-distributed_b200/csrc/pt_sphere.cuh"
#include <algorithm>
#include <climits>
#include <numeric>

using namespace b200pt;

// bounds a ray is clipped to before traversal: the same padding rule as the library (api.cu make_trav_bounds)
static TravBounds trav_bounds_of(const float *lo, const float *hi) {
    TravBounds b;
    float absmax = 0.f, ext = 0.f;
    bool ok = true;
    for (int a = 0; a < 3; ++a) ok = ok && lo[a] <= hi[a];
    for (int a = 0; a < 3; ++a) {
        const float l = ok ? lo[a] : 0.f, h = ok ? hi[a] : 0.f;
        absmax = std::max(absmax, std::max(std::fabs(l), std::fabs(h)));
        ext = std::max(ext, h - l);
    }
    const float pad = 0x1p-14f * absmax + 1e-30f;
    for (int a = 0; a < 3; ++a) {
        b.lo[a] = (ok ? lo[a] : 0.f) - pad;
        b.hi[a] = (ok ? hi[a] : 0.f) + pad;
    }
    b.scale = std::max(absmax, ext);
    return b;
}

static uint32_t bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

int main(int argc, char **argv) {
    int64_t nTris = argc > 1 ? atoll(argv[1]) : 20000;
    int64_t nRays = argc > 2 ? atoll(argv[2]) : 20000;
    unsigned seed = argc > 3 ? atoi(argv[3]) : 1;
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    float s = 0.5f * std::pow((float)nTris, -1.f / 3.f);
    std::vector<float> verts(9 * nTris);
    for (int64_t i = 0; i < nTris; ++i) {
        float c[3] = {U(rng), U(rng), U(rng)};
        for (int v = 0; v < 3; ++v)
            for (int a = 0; a < 3; ++a) verts[9 * i + 3 * v + a] = c[a] + s * U(rng);
    }
    // a few adversarial triangles: degenerate, axis-aligned flat, duplicated
    if (nTris > 16) {
        for (int a = 0; a < 9; ++a) verts[9 * 3 + a] = verts[9 * 3 + (a % 3)];          // zero area (point)
        for (int v = 0; v < 3; ++v) verts[9 * 5 + 3 * v + 2] = 0.25f;                  // flat in z
        memcpy(&verts[9 * 7], &verts[9 * 8], 36);                                      // exact duplicate
    }
    std::vector<int32_t> mat(nTris, 0), light(nTris, -1);
    b200pt_material m;
    memset(&m, 0, sizeof(m));
    m.kd[0] = m.kd[1] = m.kd[2] = .5f;
    b200pt_scene_desc sd;
    memset(&sd, 0, sizeof(sd));
    sd.n_triangles = nTris;
    sd.vertices = verts.data();
    sd.material_id = mat.data();
    sd.light_id = light.data();
    sd.n_materials = 1;
    sd.materials = &m;
    oracle_scene *os = oracle_scene_create(&sd);

    std::vector<uint8_t> degenerate(nTris);
    for (int64_t i = 0; i < nTris; ++i) {
        const float *v = &verts[9 * i];
        V3 a, b;
        TriShading sh;
        default_shading(&sh);
        degenerate[i] = !triangle_partials(mk(v[0], v[1], v[2]), mk(v[3], v[4], v[5]), mk(v[6], v[7], v[8]), sh.uv, &a, &b);
    }
    Wbvh bvh;
    build_wbvh(verts.data(), nTris, mat.data(), light.data(), nullptr, degenerate.data(), 8, &bvh);
    int64_t bad = validate_wbvh(bvh);
    const TravBounds tb = trav_bounds_of(bvh.bounds_lo, bvh.bounds_hi);
    printf("wbvh: %zu nodes, %zu tris (%u in leaves), depth %d, validation violations %lld\n", bvh.nodes.size(),
           bvh.tris.size(), bvh.n_in_leaves, bvh.max_depth, (long long)bad);
    int fail = bad != 0;

    std::vector<b200pt_ray> rays(nRays);
    for (int64_t i = 0; i < nRays; ++i) {
        b200pt_ray &r = rays[i];
        for (int a = 0; a < 3; ++a) {
            r.o[a] = 1.3f * U(rng);
            r.d[a] = U(rng);
        }
        r.t_max = (i % 3 == 0) ? 0.2f + 0.8f * std::fabs(U(rng)) : INFINITY;
        r.pad = 0;
        if (i % 17 == 0) r.d[i % 3] = 0.f;                          // axis-parallel component
        if (i % 29 == 0) { r.d[0] = r.d[1] = 0.f; r.d[2] = 1.f; }   // axis-aligned
        if (i % 31 == 0 && nTris > 0) {                             // aimed exactly at a vertex (watertightness)
            const float *v = &verts[9 * (i % nTris)];
            for (int a = 0; a < 3; ++a) r.d[a] = v[a] - r.o[a];
        }
    }
    std::vector<b200pt_hit> want(nRays);
    std::vector<uint8_t> wantOcc(nRays);
    oracle_trace_closest(os, rays.data(), want.data(), nRays);
    oracle_trace_any(os, rays.data(), wantOcc.data(), nRays);
    const U4 *nodes = reinterpret_cast<const U4 *>(bvh.nodes.data());
    const F4 *tris = reinterpret_cast<const F4 *>(bvh.tris.data());
    int64_t nHit = 0, badTri = 0, badT = 0, badOcc = 0, tieOk = 0;
    TraceCounters ctr = {0, 0};
    uint32_t overflow = 0;
    for (int64_t i = 0; i < nRays; ++i) {
        V3 o = mk(rays[i].o[0], rays[i].o[1], rays[i].o[2]), d = mk(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        TriHit h = {0, 0, 0, 0};
        uint32_t ti = traverse_wbvh<false, true>(nodes, bvh.tri_base.data(), tris, tb, nullptr, o, d, rays[i].t_max, &h, &ctr, &overflow);
        int32_t prim = ti == B200PT_MISS ? -1 : (int32_t)bvh.tris[ti].prim;
        if (prim >= 0) ++nHit;
        if (prim != want[i].triangle) {
            // an exact-t tie between two triangles may legitimately resolve differently
            if (prim >= 0 && want[i].triangle >= 0 && bits(h.t) == bits(want[i].t))
                ++tieOk;
            else
                ++badTri;
        } else if (prim >= 0 && (bits(h.t) != bits(want[i].t) || bits(h.b0) != bits(want[i].b0) ||
                                 bits(h.b1) != bits(want[i].b1)))
            ++badT;
        TriHit h2;
        uint32_t occ = traverse_wbvh<true, false>(nodes, bvh.tri_base.data(), tris, tb, nullptr, o, d, rays[i].t_max, &h2, &ctr, &overflow);
        if ((occ != B200PT_MISS) != (wantOcc[i] != 0)) ++badOcc;
    }
    printf("rays %lld: hits %lld, wrong triangle %lld (ties resolved differently: %lld), wrong t/b %lld, wrong any-hit %lld; "
           "%.2f nodes/ray %.2f tris/ray (closest)\n",
           (long long)nRays, (long long)nHit, (long long)badTri, (long long)tieOk, (long long)badT, (long long)badOcc,
           (double)ctr.nodes / nRays, (double)ctr.tris / nRays);
    if (overflow) printf("traversal stack overflows: %u\n", overflow);
    fail |= (badTri || badT || badOcc || overflow);
    // ---- the on-device builder (lbvh.cuh), emulated: same tree checks, same traversal answers
    {
        LbvhCtx c;
        memset(&c, 0, sizeof(c));
        c.vertices = verts.data();
        c.material_id = mat.data();
        c.light_id = light.data();
        c.n = nTris;
        std::vector<uint32_t> validIdx(nTris), sortedV(nTris), p2t(nTris);
        std::vector<uint64_t> keys(nTris);
        uint32_t counters[4] = {0, 1, 0, 0};
        int32_t cb[12] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
        c.valid_idx = validIdx.data();
        c.n_valid = &counters[0];
        c.n_nodes = &counters[1];
        c.n_tris = &counters[2];
        c.q_out_count = &counters[3];
        c.cbounds = cb;
        c.keys = keys.data();
        c.sorted = sortedV.data();
        c.prim_to_tri = p2t.data();
        Wbvh g;
        g.tris.resize(nTris);
        c.tris = g.tris.data();
        for (int64_t i = 0; i < nTris; ++i) lbvh_prep(c, i);
        const int64_t m = counters[0];
        c.m = m;
        g.nodes.resize((size_t)m + 16);
        g.tri_base.assign((size_t)m + 16, 0u);
        c.nodes = g.nodes.data();
        c.tri_base = g.tri_base.data();
        c.node_cap = g.nodes.size();
        {
            float absmax = 0.f;
            for (int a = 0; a < 3; ++a) {
                g.bounds_lo[a] = m ? lb_ordered_to_float(cb[6 + a]) : 0.f;
                g.bounds_hi[a] = m ? lb_ordered_to_float(cb[9 + a]) : 0.f;
                absmax = std::max(absmax, std::max(std::fabs(g.bounds_lo[a]), std::fabs(g.bounds_hi[a])));
            }
            c.cell_floor = B200PT_CELL_FLOOR * absmax;
        }
        for (int64_t k = 0; k < m; ++k) lbvh_key(c, k);
        std::vector<uint32_t> order(m);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
        std::vector<uint64_t> k2(m);
        std::vector<uint32_t> s2(m);
        for (int64_t k = 0; k < m; ++k) {
            k2[k] = keys[order[k]];
            s2[k] = sortedV[order[k]];
        }
        keys.swap(k2);
        sortedV.swap(s2);
        c.keys = keys.data();
        c.sorted = sortedV.data();
        std::vector<int32_t> L(m), R(m), P(2 * m + 1), F(m), La(m);
        std::vector<float> nbox(6 * (size_t)m + 6);
        std::vector<uint32_t> arr(m + 1);
        c.left = L.data();
        c.right = R.data();
        c.parent = P.data();
        c.first = F.data();
        c.last = La.data();
        c.nbox = nbox.data();
        c.arrivals = arr.data();
        for (int64_t i = 0; i + 1 < m; ++i) lbvh_karras(c, i);
        if (m > 1) {  // bottom-up fit in post-order
            std::vector<std::pair<int32_t, int>> st;
            st.push_back({0, 0});
            while (!st.empty()) {
                auto &top = st.back();
                if (top.second == 0) {
                    top.second = 1;
                    const int32_t n0 = top.first;
                    if (R[n0] >= 0) st.push_back({R[n0], 0});
                    if (L[n0] >= 0) st.push_back({L[n0], 0});
                } else {
                    lb_fit_node(c, top.first);
                    st.pop_back();
                }
            }
        }
        std::vector<LbvhItem> q0(m + 1), q1(m + 1);
        q0[0].node2 = m > 1 ? 0 : ~0;
        q0[0].wide = 0;
        uint32_t nItems = m > 0 ? 1 : 0;
        int depth = 1;
        while (nItems) {
            counters[3] = 0;
            c.q_in = q0.data();
            c.q_out = q1.data();
            for (uint32_t i = 0; i < nItems; ++i) lbvh_collapse(c, q0[i]);
            nItems = counters[3];
            q0.swap(q1);
            if (nItems) ++depth;
        }
        if (m == 0) {
            memset(&g.nodes[0], 0, sizeof(WbvhNode));
            const int none[B200PT_WIDTH] = {-1, -1, -1, -1, -1, -1, -1};
            wbvh_encode_node(nullptr, none, nullptr, 1.f, &g.nodes[0]);
        }
        g.n_in_leaves = counters[2];
        for (int64_t i = 0; i < nTris; ++i) lbvh_leftover(c, i);
        g.nodes.resize(counters[1]);
        g.tri_base.resize(counters[1]);
        g.max_depth = depth;
        g.prim_to_tri = p2t;
        const int64_t badG = validate_wbvh(g);
        const TravBounds gb = trav_bounds_of(g.bounds_lo, g.bounds_hi);
        const U4 *gn = reinterpret_cast<const U4 *>(g.nodes.data());
        const F4 *gt = reinterpret_cast<const F4 *>(g.tris.data());
        int64_t wrong = 0, wrongOcc = 0;
        TraceCounters gc = {0, 0};
        for (int64_t i = 0; i < nRays; ++i) {
            V3 o = mk(rays[i].o[0], rays[i].o[1], rays[i].o[2]), d = mk(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
            TriHit h = {0, 0, 0, 0};
            uint32_t ti = traverse_wbvh<false, true>(gn, g.tri_base.data(), gt, gb, nullptr, o, d, rays[i].t_max, &h, &gc, &overflow);
            int32_t prim = ti == B200PT_MISS ? -1 : (int32_t)g.tris[ti].prim;
            const bool tie = prim >= 0 && want[i].triangle >= 0 && bits(h.t) == bits(want[i].t);
            if (prim != want[i].triangle && !tie) ++wrong;
            if (prim == want[i].triangle && prim >= 0 && bits(h.t) != bits(want[i].t)) ++wrong;
            TriHit h2;
            uint32_t occ = traverse_wbvh<true, false>(gn, g.tri_base.data(), gt, gb, nullptr, o, d, rays[i].t_max, &h2, &gc, &overflow);
            if ((occ != B200PT_MISS) != (wantOcc[i] != 0)) ++wrongOcc;
        }
        printf("lbvh builder (emulated): %lld of %lld triangles in the tree, %zu nodes, depth %d, violations %lld, wrong hits %lld, "
               "wrong any-hit %lld, %.2f nodes/ray %.2f tris/ray (host SAH tree: %.2f, %.2f)\n",
               (long long)m, (long long)nTris, g.nodes.size(), depth, (long long)badG, (long long)wrong, (long long)wrongOcc,
               (double)gc.nodes / nRays, (double)gc.tris / nRays, (double)ctr.nodes / nRays, (double)ctr.tris / nRays);
        fail |= (badG || wrong || wrongOcc || overflow || depth > B200PT_STACK - 4);
    }
    oracle_scene_destroy(os);

    // ---- spheres: the device routine (compiled for the host) against the oracle's Scene::Intersect
    {
        const int nSph = 6;
        std::vector<b200pt_sphere> sph(nSph);
        std::vector<DevSphere> dev(nSph);
        for (int k = 0; k < nSph; ++k) {
            b200pt_sphere &sp = sph[k];
            memset(&sp, 0, sizeof(sp));
            const float c[3] = {U(rng), U(rng), U(rng)};
            const float sc[3] = {k % 2 ? 0.6f + 0.5f * std::fabs(U(rng)) : 1.f, k % 2 ? 0.6f + 0.5f * std::fabs(U(rng)) : 1.f,
                                 k % 3 == 2 ? -1.2f : 1.f};
            for (int a = 0; a < 4; ++a) sp.object_to_world[5 * a] = sp.world_to_object[5 * a] = 1.f;
            for (int a = 0; a < 3; ++a) {  // Translate(c) * Scale(sc): m = T S, mInv = S^-1 T^-1
                sp.object_to_world[5 * a] = sc[a];
                sp.object_to_world[4 * a + 3] = c[a];
                sp.world_to_object[5 * a] = 1.f / sc[a];
                sp.world_to_object[4 * a + 3] = (1.f / sc[a]) * -c[a];
            }
            sp.radius = 0.15f + 0.3f * std::fabs(U(rng));
            sp.material_id = 0;
            sp.light_id = -1;
            sp.reverse_orientation = k == 1;
            sp.transform_swaps_handedness = sc[2] < 0;
            DevSphere &d = dev[k];
            memcpy(d.o2w, sp.object_to_world, 64);
            memcpy(d.w2o, sp.world_to_object, 64);
            d.radius = sp.radius;
            d.mat_flags = 0;
            d.z_min = -sp.radius;
            d.z_max = sp.radius;
            d.phi_max = (PT_PI / 180) * 360.f;
            d.theta_min = pt_acosf(-1.f);
            d.theta_max = pt_acosf(1.f);
            if (k >= 3) {  // clipped spheres: the members the Sphere ctor would compute
                float zp[5];
                const float zmin = -0.5f * sp.radius, zmax = (k == 4 ? 0.7f : 1.f) * sp.radius, phimax = k == 5 ? 200.f : 360.f;
                zp[0] = pt_clamp(pt_min(zmin, zmax), -sp.radius, sp.radius);
                zp[1] = pt_clamp(pt_max(zmin, zmax), -sp.radius, sp.radius);
                zp[2] = pt_acosf(pt_clamp(pt_min(zmin, zmax) / sp.radius, -1.f, 1.f));
                zp[3] = pt_acosf(pt_clamp(pt_max(zmin, zmax) / sp.radius, -1.f, 1.f));
                zp[4] = (PT_PI / 180) * pt_clamp(phimax, 0.f, 360.f);
                sp.z_min = d.z_min = zp[0];
                sp.z_max = d.z_max = zp[1];
                sp.theta_min = d.theta_min = zp[2];
                sp.theta_max = d.theta_max = zp[3];
                sp.phi_max = d.phi_max = zp[4];
            }
            // Shape::WorldBound(): what the oracle and the library use when leaf_bounds is left zero
            for (int a = 0; a < 3; ++a) {
                d.leaf_lo[a] = INFINITY;
                d.leaf_hi[a] = -INFINITY;
            }
            for (int cn = 0; cn < 8; ++cn) {
                const float r = sp.radius;
                const V3 q = xform_point(d.o2w, mk((cn & 1) ? r : -r, (cn & 2) ? r : -r, (cn & 4) ? d.z_max : d.z_min));
                for (int a = 0; a < 3; ++a) {
                    d.leaf_lo[a] = std::min(d.leaf_lo[a], comp(q, a));
                    d.leaf_hi[a] = std::max(d.leaf_hi[a], comp(q, a));
                }
            }
        }
        b200pt_scene_desc sd2;
        memset(&sd2, 0, sizeof(sd2));
        sd2.n_materials = 1;
        sd2.materials = &m;
        sd2.n_spheres = nSph;
        sd2.spheres = sph.data();
        oracle_scene *os2 = oracle_scene_create(&sd2);
        std::vector<b200pt_hit> w2(nRays);
        std::vector<uint8_t> o2(nRays);
        for (int64_t i = 0; i < nRays; ++i)
            if (i % 5 == 0)  // start some rays inside a sphere
                for (int a = 0; a < 3; ++a) rays[i].o[a] = sph[i % nSph].object_to_world[4 * a + 3] + 0.05f * U(rng);
        oracle_trace_closest(os2, rays.data(), w2.data(), nRays);
        oracle_trace_any(os2, rays.data(), o2.data(), nRays);
        int64_t nh = 0, badS = 0, badO = 0;
        for (int64_t i = 0; i < nRays; ++i) {
            V3 o = mk(rays[i].o[0], rays[i].o[1], rays[i].o[2]), d = mk(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
            float tmax = rays[i].t_max, t;
            int best = -1;
            bool occ = false;
            for (int k = 0; k < nSph; ++k) {
                if (sphere_leaf_test(dev[k], o, d, rays[i].t_max) && sphere_intersect(dev[k], o, d, rays[i].t_max, &t, nullptr))
                    occ = true;
                if (sphere_leaf_test(dev[k], o, d, tmax) && sphere_intersect(dev[k], o, d, tmax, &t, nullptr)) {
                    tmax = t;
                    best = k;
                }
            }
            if (best >= 0) ++nh;
            if (best != w2[i].triangle || (best >= 0 && bits(tmax) != bits(w2[i].t))) ++badS;
            if (occ != (o2[i] != 0)) ++badO;
        }
        printf("spheres: %lld rays, %lld hits, wrong sphere/t %lld, wrong any-hit %lld\n", (long long)nRays, (long long)nh,
               (long long)badS, (long long)badO);
        fail |= (badS || badO);
        // Sphere::Sample(ref, u) / Sphere::Pdf(ref, wi): reference points outside (cone sampling) and inside (area sampling,
        // Shape::Pdf through Intersect) every sphere
        int64_t badSample = 0, badPdf = 0, nInside = 0;
        for (int64_t i = 0; i < 20000; ++i) {
            const int k = (int)(i % nSph);
            DevSphere d = dev[k];
            d.area = d.phi_max * d.radius * (d.z_max - d.z_min);
            d.reverse_orientation = sph[k].reverse_orientation;
            const bool inside = (i % 3) == 0;
            const float rr = inside ? 0.6f * sph[k].radius * std::fabs(U(rng)) : 2.f + 3.f * std::fabs(U(rng));
            V3 dir = normalize(mk(U(rng), U(rng), U(rng) + 1e-3f));
            const V3 ctr = mk(sph[k].object_to_world[3], sph[k].object_to_world[7], sph[k].object_to_world[11]);
            const V3 refP = ctr + dir * rr, refN = normalize(mk(U(rng), U(rng), U(rng) + 1e-3f));
            const V3 refE = mk(1e-6f * std::fabs(U(rng)), 1e-6f * std::fabs(U(rng)), 1e-6f * std::fabs(U(rng)));
            const float u[2] = {0.5f * (U(rng) + 1.f) * 0.999f, 0.5f * (U(rng) + 1.f) * 0.999f};
            float pdf = 0, want10[10];
            const LightSample ls = sphere_sample(d, refP, refE, refN, u, &pdf);
            const float rp[3] = {refP.x, refP.y, refP.z}, re[3] = {refE.x, refE.y, refE.z}, rn[3] = {refN.x, refN.y, refN.z};
            oracle_sphere_sample(&sph[k], rp, re, rn, u, want10);
            const float got10[10] = {ls.p.x, ls.p.y, ls.p.z, ls.n.x, ls.n.y, ls.n.z, ls.pError.x, ls.pError.y, ls.pError.z, pdf};
            for (int a = 0; a < 10; ++a)
                if (bits(got10[a]) != bits(want10[a]) && !(got10[a] != got10[a] && want10[a] != want10[a])) {
                    ++badSample;
                    break;
                }
            const V3 wi = normalize(ls.p - refP);
            const float wv[3] = {wi.x, wi.y, wi.z};
            const float pg = sphere_pdf(d, refP, refE, refN, wi), pw = oracle_sphere_pdf(&sph[k], rp, re, rn, wv);
            if (bits(pg) != bits(pw) && !(pg != pg && pw != pw)) ++badPdf;
            nInside += inside;
        }
        printf("sphere sampling: 20000 reference points (%lld inside), wrong samples %lld, wrong pdfs %lld\n", (long long)nInside,
               (long long)badSample, (long long)badPdf);
        fail |= (badSample || badPdf);
        oracle_scene_destroy(os2);
        long badA = 0;
        for (uint32_t u = 0; u <= 0x3f800000u; u += 3)
            for (int sg = 0; sg < 2; ++sg) {
                float x;
                uint32_t ub = u | (sg ? 0x80000000u : 0u);
                memcpy(&x, &ub, 4);
                if (bits(std::acos(x)) != bits(pt_acosf(x))) ++badA;
            }
        printf("acosf: %ld mismatches against the host libm over [-1, 1] (every third float)\n", badA);
        fail |= badA != 0;
    }
    // ---- object instances: the device routines (compiled for the host) against the oracle's TransformedPrimitive
    if (nTris >= 50) {
        const int64_t nObj = std::min<int64_t>(nTris, 3000);
        std::vector<float> ov(verts.begin(), verts.begin() + 9 * nObj);
        for (float &v : ov) v *= 0.3f;
        Wbvh ob;
        std::vector<uint8_t> degO(degenerate.begin(), degenerate.begin() + nObj);
        build_wbvh(ov.data(), nObj, mat.data(), light.data(), nullptr, degO.data(), 4, &ob);
        const TravBounds obb = trav_bounds_of(ob.bounds_lo, ob.bounds_hi);
        const int nInst = 4;
        std::vector<b200pt_instance> insts(nInst);
        std::vector<DevInstance> dev(nInst);
        for (int k = 0; k < nInst; ++k) {
            b200pt_instance &in = insts[k];
            memset(&in, 0, sizeof(in));
            in.first_triangle = 0;
            in.n_triangles = nObj;
            const float c[3] = {k == 0 ? 0.f : U(rng), k == 0 ? 0.f : U(rng), k == 0 ? 0.f : U(rng)};
            const float sc[3] = {k == 2 ? 1.4f : 1.f, k == 2 ? 0.7f : 1.f, k == 3 ? -1.2f : 1.f};
            for (int a = 0; a < 4; ++a) in.instance_to_world[5 * a] = in.world_to_instance[5 * a] = 1.f;
            for (int a = 0; a < 3; ++a) {
                in.instance_to_world[5 * a] = sc[a];
                in.instance_to_world[4 * a + 3] = c[a];
                in.world_to_instance[5 * a] = 1.f / sc[a];
                in.world_to_instance[4 * a + 3] = (1.f / sc[a]) * -c[a];
            }
            in.is_identity = k == 0;
            DevInstance &d = dev[k];
            memcpy(d.i2w, in.instance_to_world, 64);
            memcpy(d.w2i, in.world_to_instance, 64);
            d.is_identity = in.is_identity;
            d.node_off = d.tri_off = 0;
            // the instance's own WorldBound(): what the oracle and the library use when leaf_bounds is zero
            float olo[3] = {INFINITY, INFINITY, INFINITY}, ohi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int64_t v = 0; v < 3 * nObj; ++v)
                for (int a = 0; a < 3; ++a) {
                    olo[a] = std::min(olo[a], ov[3 * v + a]);
                    ohi[a] = std::max(ohi[a], ov[3 * v + a]);
                }
            for (int a = 0; a < 3; ++a) {
                d.leaf_lo[a] = INFINITY;
                d.leaf_hi[a] = -INFINITY;
            }
            for (int cn = 0; cn < 8; ++cn) {
                const V3 q = xform_point(d.i2w, mk((cn & 1) ? ohi[0] : olo[0], (cn & 2) ? ohi[1] : olo[1], (cn & 4) ? ohi[2] : olo[2]));
                for (int a = 0; a < 3; ++a) {
                    d.leaf_lo[a] = std::min(d.leaf_lo[a], comp(q, a));
                    d.leaf_hi[a] = std::max(d.leaf_hi[a], comp(q, a));
                }
            }
        }
        b200pt_scene_desc sd3;
        memset(&sd3, 0, sizeof(sd3));
        sd3.n_triangles = nObj;
        sd3.vertices = ov.data();
        sd3.material_id = mat.data();
        sd3.light_id = light.data();
        sd3.n_materials = 1;
        sd3.materials = &m;
        sd3.n_instances = nInst;
        sd3.instances = insts.data();
        sd3.n_toplevel_triangles = 0;
        oracle_scene *os3 = oracle_scene_create(&sd3);
        std::vector<b200pt_hit> w3(nRays);
        std::vector<uint8_t> o3(nRays);
        oracle_trace_closest(os3, rays.data(), w3.data(), nRays);
        oracle_trace_any(os3, rays.data(), o3.data(), nRays);
        const U4 *on = reinterpret_cast<const U4 *>(ob.nodes.data());
        const F4 *ot = reinterpret_cast<const F4 *>(ob.tris.data());
        int64_t nh = 0, badI = 0, badO = 0, ties = 0;
        TraceCounters ic = {0, 0};
        for (int64_t i = 0; i < nRays; ++i) {
            V3 o = mk(rays[i].o[0], rays[i].o[1], rays[i].o[2]), d = mk(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
            float tmax = rays[i].t_max;
            int32_t prim = -1;
            float tBest = 0;
            bool occ = false;
            for (int k = 0; k < nInst; ++k) {
                V3 o2, d2;
                float tm2;
                TriHit h;
                if (instance_leaf_test(dev[k], o, d, rays[i].t_max)) {
                    instance_ray(dev[k], o, d, rays[i].t_max, &o2, &d2, &tm2);
                    if (traverse_wbvh<true, false>(on, ob.tri_base.data(), ot, obb, nullptr, o2, d2, tm2, &h, &ic, &overflow) != B200PT_MISS) occ = true;
                }
                if (getenv("PREFLIGHT_VERBOSE") && (i == 4505 || i == 15674)) {
                    instance_ray(dev[k], o, d, tmax, &o2, &d2, &tm2);
                    TriHit hb = {0, 0, 0, 0};
                    uint32_t tb = traverse_wbvh<false, false>(on, ob.tri_base.data(), ot, obb, nullptr, o2, d2, tm2, &hb, &ic, &overflow);
                    printf("   ray %lld inst %d: leaf %d o2=(%g %g %g) d2=(%g %g %g) tm2=%g hit %u t=%g\n", (long long)i, k,
                           (int)instance_leaf_test(dev[k], o, d, tmax), o2.x, o2.y, o2.z, d2.x, d2.y, d2.z, tm2, tb, hb.t);
                }
                if (!instance_leaf_test(dev[k], o, d, tmax)) continue;
                instance_ray(dev[k], o, d, tmax, &o2, &d2, &tm2);
                const uint32_t ti = traverse_wbvh<false, false>(on, ob.tri_base.data(), ot, obb, nullptr, o2, d2, tm2, &h, &ic, &overflow);
                if (ti == B200PT_MISS) continue;
                tmax = h.t;
                tBest = h.t;
                prim = (int32_t)ob.tris[ti].prim;
            }
            if (prim >= 0) ++nh;
            if (prim != w3[i].triangle) {
                if (prim >= 0 && w3[i].triangle >= 0 && bits(tBest) == bits(w3[i].t))
                    ++ties;
                else {
                    ++badI;
                    if (getenv("PREFLIGHT_VERBOSE")) printf("  ray %lld: prim %d t %a vs oracle %d t %a tmax %a\n", (long long)i, prim, tBest, w3[i].triangle, w3[i].t, rays[i].t_max);
                }
            } else if (prim >= 0 && bits(tBest) != bits(w3[i].t))
                ++badI;
            if (occ != (o3[i] != 0)) {
                ++badO;
                if (getenv("PREFLIGHT_VERBOSE")) printf("  ray %lld: occ %d vs oracle %d tmax %a\n", (long long)i, (int)occ, (int)o3[i], rays[i].t_max);
            }
        }
        printf("instances: %lld rays, %lld hits, wrong triangle/t %lld (ties %lld), wrong any-hit %lld\n", (long long)nRays,
               (long long)nh, (long long)badI, (long long)ties, (long long)badO);
        fail |= (badI || badO);
        oracle_scene_destroy(os3);
    }
    printf(fail ? "PREFLIGHT FAILED\n" : "PREFLIGHT OK\n");
    return fail;
}
