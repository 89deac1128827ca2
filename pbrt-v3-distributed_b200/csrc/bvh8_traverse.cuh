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

namespace B200PT_NS {

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

// Byte j of w spliced into the mantissa of 2^23: the float 2^23 + byte, exactly, with one PRMT and
// no integer->float conversion.  The node test folds the "- 2^23" into the FMA's addend.
// `magic` must hold 0x4B000000 in a REGISTER (the kernel passes it as a run-time argument): PRMT takes
// one immediate, and it should be the byte selector -- otherwise the compiler re-materialises a
// selector register for every one of the 48 extractions of a node.
#ifdef __CUDA_ARCH__
template <int J>
B200_D float byte_plus_2p23(uint32_t w, uint32_t magic) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(magic), "n"(0x7650 | J));
    return __uint_as_float(r);
}
#else
template <int J>
inline float byte_plus_2p23(uint32_t w, uint32_t) { return 8388608.0f + (float)((w >> (8 * J)) & 0xffu); }
#endif

// Per-ray traversal state.  One step = "take the next child group: fetch its node
// and test the eight children, then intersect the leaf triangles that were hit".
struct Trav {
    V3 o;
    RayShear sh;
    float idx, idy, idz;
    uint32_t oct, octinv;
    float tmax;
    uint32_t best;
    TriHit hit;
    uint32_t cur_x, cur_y;
    int sp;
    uint32_t magic;  // 0x4B000000 kept in a register, see byte_plus_2p23
};
// The stack of postponed child groups lives in its own object so that the scalar
// state above stays in registers (a struct with a dynamically indexed array is
// placed in local memory as a whole).
struct TravStack {
    uint32_t x[B200PT_STACK], y[B200PT_STACK];
};

B200_HD void trav_init(Trav &T, const V3 &o, const V3 &d, float rayTMax) {
    T.o = o;
    T.sh = make_shear(d);
    T.idx = safe_rcp_dir(d.x);
    T.idy = safe_rcp_dir(d.y);
    T.idz = safe_rcp_dir(d.z);
    // the octant follows the sign BIT, like safe_rcp_dir does: a component of -0.0 (mirrored instances, reflections)
    // must pick the same near / far planes as the sign of its reciprocal
    T.oct = (float_as_uint(d.x) >> 31) | ((float_as_uint(d.y) >> 31) << 1) | ((float_as_uint(d.z) >> 31) << 2);
    T.octinv = 7u - T.oct;
    T.tmax = rayTMax;
    T.best = B200PT_MISS;
    T.hit.t = T.hit.b0 = T.hit.b1 = T.hit.b2 = 0.f;
    T.cur_x = 0u;
    T.cur_y = 0x80000000u;  // the root as a one-child group
    T.sp = 0;
    T.magic = 0x4B000000u;
}

// Node phase: take the next inner child of the current group, fetch its node, test the eight
// children.  Leaves the hit inner children in T.cur and returns the hit leaf triangles as a
// triangle group (*tg_x = first triangle, *tg_y = bit per triangle).  Requires T.cur to be a node group.
template <bool COUNT>
B200_HD void trav_node_phase(Trav &T, TravStack &S, const U4 *__restrict__ nodes, uint32_t *tg_x, uint32_t *tg_y,
                             TraceCounters *ctr) {
        const uint32_t hits = T.cur_y;
        const int bit = msb32(hits);
        T.cur_y &= ~(1u << bit);
        if (T.cur_y & 0xff000000u) {
            if (T.sp < B200PT_STACK) {
                S.x[T.sp] = T.cur_x;
                S.y[T.sp] = T.cur_y;
                ++T.sp;
            }
        }
        const uint32_t slot = ((uint32_t)(bit - 24)) ^ T.octinv;
        const uint32_t rel = (uint32_t)popc32(hits & 0xffu & ((1u << slot) - 1u));
        const U4 *np = nodes + (size_t)(T.cur_x + rel) * 5;
        const U4 n0 = ld_u4(np), n1 = ld_u4(np + 1), n2 = ld_u4(np + 2), n3 = ld_u4(np + 3), n4 = ld_u4(np + 4);
        if (COUNT) ctr->nodes++;
        // n0: p.xyz | e.x e.y e.z imask      n1: child_base tri_base meta[0..3] meta[4..7]
        // n2: qlo.x[0..7] qlo.y[0..7]        n3: qlo.z[0..7] qhi.x[0..7]      n4: qhi.y[0..7] qhi.z[0..7]
        // t(q) = (p + q*scale - o) / d = q*a + b with a = scale/d, b = (p - o)/d.  The byte arrives as
        // 2^23 + q, so the FMA uses the addend c = b - 2^23*a; the product is exact inside the FMA and
        // the only extra rounding (in c) is below half a grid cell -- the builder leaves one full cell of slack.
        const float ax = uint_as_float((n0.w & 0xffu) << 23) * T.idx;
        const float ay = uint_as_float(((n0.w >> 8) & 0xffu) << 23) * T.idy;
        const float az = uint_as_float(((n0.w >> 16) & 0xffu) << 23) * T.idz;
        const float cx = fma_any(-8388608.0f, ax, (uint_as_float(n0.x) - T.o.x) * T.idx);
        const float cy = fma_any(-8388608.0f, ay, (uint_as_float(n0.y) - T.o.y) * T.idy);
        const float cz = fma_any(-8388608.0f, az, (uint_as_float(n0.z) - T.o.z) * T.idz);
        const uint32_t imask = n0.w >> 24;
        const bool nxg = (T.oct & 1u) != 0, nyg = (T.oct & 2u) != 0, nzg = (T.oct & 4u) != 0;
        const uint32_t octinv4 = T.octinv * 0x01010101u;
        uint32_t hitmask = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // near / far quantised planes per axis according to the ray's direction sign
            const uint32_t qlox = h ? n2.y : n2.x, qloy = h ? n2.w : n2.z, qloz = h ? n3.y : n3.x;
            const uint32_t qhix = h ? n3.w : n3.z, qhiy = h ? n4.y : n4.x, qhiz = h ? n4.w : n4.z;
            const uint32_t nx = nxg ? qhix : qlox, fx = nxg ? qlox : qhix;
            const uint32_t ny = nyg ? qhiy : qloy, fy = nyg ? qloy : qhiy;
            const uint32_t nz = nzg ? qhiz : qloz, fz = nzg ? qloz : qhiz;
            // four children at a time (Ylitie et al.): inner children (meta = 001sssss, sssss = 24 + slot)
            // get the bit 24 + (slot ^ octinv); leaves get their unary triangle count at their offset
            const uint32_t meta4 = h ? n1.w : n1.z;
            const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
            const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xffu;
            const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
            const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
#define B200PT_CHILD(J)                                                                                              \
    {                                                                                                                \
        const float tn = fmaxf(fmaxf(fma_any(byte_plus_2p23<J>(nx, T.magic), ax, cx),                                \
                                     fma_any(byte_plus_2p23<J>(ny, T.magic), ay, cy)),                               \
                               fmaxf(fma_any(byte_plus_2p23<J>(nz, T.magic), az, cz), 0.f));                         \
        const float tf = fminf(fminf(fma_any(byte_plus_2p23<J>(fx, T.magic), ax, cx),                                \
                                     fma_any(byte_plus_2p23<J>(fy, T.magic), ay, cy)),                               \
                               fminf(fma_any(byte_plus_2p23<J>(fz, T.magic), az, cz), T.tmax));                      \
        const uint32_t bits = ((child_bits4 >> (8 * J)) & 0xffu) << ((bit_index4 >> (8 * J)) & 0xffu);               \
        hitmask |= (tn <= tf) ? bits : 0u;                                                                           \
    }
            B200PT_CHILD(0)
            B200PT_CHILD(1)
            B200PT_CHILD(2)
            B200PT_CHILD(3)
#undef B200PT_CHILD
        }
        T.cur_x = n1.x;
        T.cur_y = (hitmask & 0xff000000u) | imask;
        *tg_x = n1.y;
        *tg_y = hitmask & 0x00ffffffu;
}

// Triangle phase: exact watertight tests of a triangle group.  Returns true if ANY_HIT found a hit.
template <bool ANY_HIT, bool COUNT>
B200_HD bool trav_tri_phase(Trav &T, const F4 *__restrict__ tris, uint32_t tg_x, uint32_t tg_y, TraceCounters *ctr) {
    while (tg_y) {
        const int j = msb32(tg_y);
        tg_y &= ~(1u << j);
        const uint32_t ti = tg_x + (uint32_t)j;
        const F4 *tp = tris + (size_t)ti * 3;
        const F4 v0 = ld_f4(tp), v1 = ld_f4(tp + 1), v2 = ld_f4(tp + 2);
        if (COUNT) ctr->tris++;
        TriHit h;
        if (triangle_test(mk(v0.x, v0.y, v0.z), mk(v1.x, v1.y, v1.z), mk(v2.x, v2.y, v2.z), T.o, T.sh, T.tmax, &h)) {
            T.tmax = h.t;  // primitive.cpp:120
            T.best = ti;
            T.hit = h;
            if (ANY_HIT) return true;
        }
    }
    return false;
}

// Pops the next node group if the current one is exhausted; false when nothing is left.
B200_HD bool trav_next_group(Trav &T, TravStack &S) {
    if ((T.cur_y & 0xff000000u) == 0) {
        if (T.sp == 0) return false;
        --T.sp;
        T.cur_x = S.x[T.sp];
        T.cur_y = S.y[T.sp];
    }
    return true;
}

// One complete step (node phase, then its triangles at once).  Returns true when the traversal
// is complete (closest hit known / any hit found / nothing left).
template <bool ANY_HIT, bool COUNT>
B200_HD bool trav_step(Trav &T, TravStack &S, const U4 *__restrict__ nodes, const F4 *__restrict__ tris,
                        TraceCounters *ctr) {
    uint32_t tg_x = 0, tg_y = 0;
    if (T.cur_y & 0xff000000u) trav_node_phase<COUNT>(T, S, nodes, &tg_x, &tg_y, ctr);
    if (trav_tri_phase<ANY_HIT, COUNT>(T, tris, tg_x, tg_y, ctr)) return true;
    return !trav_next_group(T, S);
}

// Whole traversal of one ray (used by the CPU pre-flight and the simple entry points).
// Returns the leaf-order index of the closest (ANY_HIT: of some) hit triangle or
// B200PT_MISS; *hit receives (t, b0, b1, b2) of the accepted intersection.
template <bool ANY_HIT, bool COUNT>
B200_HD uint32_t traverse_bvh8(const U4 *__restrict__ nodes, const F4 *__restrict__ tris, const V3 &o, const V3 &d,
                               float rayTMax, TriHit *hit, TraceCounters *ctr) {
    Trav T;
    TravStack S;
    trav_init(T, o, d, rayTMax);
    while (!trav_step<ANY_HIT, COUNT>(T, S, nodes, tris, ctr)) {
    }
    *hit = T.hit;
    return T.best;
}

}  // namespace B200PT_NS
#endif
