// bvh8_traverse.cuh -- stack-based traversal of the 8-wide compressed BVH
// (bvh8.h), one ray per thread.  Replaces BVHAccel::Intersect / IntersectP
// (accelerators/bvh.cpp:662-738) + GeometricPrimitive::Intersect
// (core/primitive.cpp:116-130) + Triangle::Intersect / IntersectP
// (shapes/triangle.cpp:188-291, 427-517).
//
// What must match the reference bit for bit is the *result*: the hit triangle
// and (t, b0, b1, b2) come from the same watertight test with the same
// shrinking ray.tMax semantics (accept tScaled == tMax*det, primitive.cpp:120).
// The box tests only have to be conservative, so they run on the quantised
// grid with fused multiply-adds.
//
// Traversal state follows Ylitie et al. 2017: the stack holds "groups" --
// (child_base, hit bits | imask) for inner children still to visit and
// (tri_base, triangle bits) for leaf triangles -- and children are visited in
// the order (slot XOR ray octant), highest first, which the builder's slot
// assignment turns into an approximate front-to-back order.
#ifndef B200PT_BVH8_TRAVERSE_CUH
#define B200PT_BVH8_TRAVERSE_CUH

#include "pt_core.cuh"

namespace b200pt {

struct U4 {
    uint32_t x, y, z, w;
};
struct F4 {
    float x, y, z, w;
};

#ifdef __CUDA_ARCH__
B200_D U4 ld_u4(const U4 *p) {
    uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    U4 r;
    r.x = v.x;
    r.y = v.y;
    r.z = v.z;
    r.w = v.w;
    return r;
}
B200_D F4 ld_f4(const F4 *p) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    F4 r;
    r.x = v.x;
    r.y = v.y;
    r.z = v.z;
    r.w = v.w;
    return r;
}
B200_D float fma_any(float a, float b, float c) { return __fmaf_rn(a, b, c); }
B200_D int msb32(uint32_t v) { return 31 - __clz((int)v); }
B200_D int popc32(uint32_t v) { return __popc(v); }
#else
inline U4 ld_u4(const U4 *p) { return *p; }
inline F4 ld_f4(const F4 *p) { return *p; }
inline float fma_any(float a, float b, float c) { return a * b + c; }
inline int msb32(uint32_t v) { return 31 - __builtin_clz(v); }
inline int popc32(uint32_t v) { return __builtin_popcount(v); }
#endif

#define B200PT_STACK 48
#define B200PT_MISS 0xffffffffu

struct TraceCounters {
    uint32_t nodes, tris;
};

B200_HD float safe_rcp_dir(float d) {
    // keeps the sign of d, avoids inf/NaN in the slab arithmetic for axis-parallel rays
    float a = pt_abs(d);
    if (!(a > 1e-20f)) d = (float_as_uint(d) & 0x80000000u) ? -1e-20f : 1e-20f;
    return 1.0f / d;
}

// Returns the leaf-order index of the closest (ANY_HIT: of some) hit triangle
// or B200PT_MISS.  *hit receives (t, b0, b1, b2) of the accepted intersection.
template <bool ANY_HIT, bool COUNT>
B200_HD uint32_t traverse_bvh8(const U4 *__restrict__ nodes, const F4 *__restrict__ tris, const V3 &o, const V3 &d,
                               float rayTMax, TriHit *hit, TraceCounters *ctr) {
    const RayShear sh = make_shear(d);
    const float idx = safe_rcp_dir(d.x), idy = safe_rcp_dir(d.y), idz = safe_rcp_dir(d.z);
    const uint32_t oct = (d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u);
    const uint32_t octinv = 7u - oct;
    float tmax = rayTMax;
    uint32_t best = B200PT_MISS;

    uint32_t stk_x[B200PT_STACK], stk_y[B200PT_STACK];
    int sp = 0;
    uint32_t cur_x = 0u, cur_y = 0x80000000u;  // the root as a one-child group

    while (true) {
        uint32_t tg_x, tg_y;
        if (cur_y & 0xff000000u) {
            const uint32_t hits = cur_y;
            const int bit = msb32(hits);
            cur_y &= ~(1u << bit);
            if (cur_y & 0xff000000u) {
                if (sp < B200PT_STACK) {
                    stk_x[sp] = cur_x;
                    stk_y[sp] = cur_y;
                    ++sp;
                }
            }
            const uint32_t slot = ((uint32_t)(bit - 24)) ^ octinv;
            const uint32_t rel = (uint32_t)popc32(hits & 0xffu & ((1u << slot) - 1u));
            const U4 *np = nodes + (size_t)(cur_x + rel) * 5;
            const U4 n0 = ld_u4(np), n1 = ld_u4(np + 1), n2 = ld_u4(np + 2), n3 = ld_u4(np + 3), n4 = ld_u4(np + 4);
            if (COUNT) ctr->nodes++;
            // n0: p.x p.y p.z (e.x e.y e.z imask) ; n1: child_base tri_base meta[0..3] meta[4..7]
            // n2: qlo.x[0..7] qlo.y[0..3] qlo.y[4..7] -> (x: qlox 0-3, y: qlox 4-7, z: qloy 0-3, w: qloy 4-7)
            // n3: qloz 0-3, qloz 4-7, qhix 0-3, qhix 4-7 ; n4: qhiy 0-3, qhiy 4-7, qhiz 0-3, qhiz 4-7
            const float px = uint_as_float(n0.x), py = uint_as_float(n0.y), pz = uint_as_float(n0.z);
            const float sx = uint_as_float((n0.w & 0xffu) << 23), sy = uint_as_float(((n0.w >> 8) & 0xffu) << 23),
                        sz = uint_as_float(((n0.w >> 16) & 0xffu) << 23);
            const uint32_t imask = n0.w >> 24;
            const float ax = sx * idx, ay = sy * idy, az = sz * idz;
            const float bx = (px - o.x) * idx, by = (py - o.y) * idy, bz = (pz - o.z) * idz;
            // near / far quantised planes per axis according to the ray's direction sign
            const uint32_t qlox[2] = {n2.x, n2.y}, qloy[2] = {n2.z, n2.w}, qloz[2] = {n3.x, n3.y};
            const uint32_t qhix[2] = {n3.z, n3.w}, qhiy[2] = {n4.x, n4.y}, qhiz[2] = {n4.z, n4.w};
            const uint32_t meta[2] = {n1.z, n1.w};
            uint32_t hitmask = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t nx = (oct & 1u) ? qhix[h] : qlox[h], fx = (oct & 1u) ? qlox[h] : qhix[h];
                const uint32_t ny = (oct & 2u) ? qhiy[h] : qloy[h], fy = (oct & 2u) ? qloy[h] : qhiy[h];
                const uint32_t nz = (oct & 4u) ? qhiz[h] : qloz[h], fz = (oct & 4u) ? qloz[h] : qhiz[h];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t m = (meta[h] >> (8 * j)) & 0xffu;
                    if (m == 0) continue;
                    const float tnx = fma_any((float)((nx >> (8 * j)) & 0xffu), ax, bx);
                    const float tny = fma_any((float)((ny >> (8 * j)) & 0xffu), ay, by);
                    const float tnz = fma_any((float)((nz >> (8 * j)) & 0xffu), az, bz);
                    const float tfx = fma_any((float)((fx >> (8 * j)) & 0xffu), ax, bx);
                    const float tfy = fma_any((float)((fy >> (8 * j)) & 0xffu), ay, by);
                    const float tfz = fma_any((float)((fz >> (8 * j)) & 0xffu), az, bz);
                    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.f));
                    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
                    if (tn <= tf) {
                        const uint32_t s = (uint32_t)(4 * h + j);
                        if (imask & (1u << s))
                            hitmask |= 1u << (24u + (s ^ octinv));
                        else
                            hitmask |= (m >> 5) << (m & 31u);
                    }
                }
            }
            cur_x = n1.x;
            cur_y = (hitmask & 0xff000000u) | imask;
            tg_x = n1.y;
            tg_y = hitmask & 0x00ffffffu;
        } else {
            tg_x = cur_x;
            tg_y = cur_y;
            cur_x = 0;
            cur_y = 0;
        }

        while (tg_y) {
            const int j = msb32(tg_y);
            tg_y &= ~(1u << j);
            const uint32_t ti = tg_x + (uint32_t)j;
            const F4 *tp = tris + (size_t)ti * 3;
            const F4 v0 = ld_f4(tp), v1 = ld_f4(tp + 1), v2 = ld_f4(tp + 2);
            if (COUNT) ctr->tris++;
            TriHit h;
            if (triangle_test(mk(v0.x, v0.y, v0.z), mk(v1.x, v1.y, v1.z), mk(v2.x, v2.y, v2.z), o, sh, tmax, &h)) {
                tmax = h.t;  // primitive.cpp:120
                best = ti;
                *hit = h;
                if (ANY_HIT) return best;
            }
        }

        if ((cur_y & 0xff000000u) == 0) {
            if (sp == 0) break;
            --sp;
            cur_x = stk_x[sp];
            cur_y = stk_y[sp];
        }
    }
    return best;
}

}  // namespace b200pt
#endif
