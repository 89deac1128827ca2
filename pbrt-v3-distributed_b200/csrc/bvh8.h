// bvh8.h -- the acceleration structure that replaces the reference's binary
// LinearBVHNode array (accelerators/bvh.cpp:95-104): an 8-wide BVH with
// quantised child boxes, 80 bytes per node (five 16-byte loads), after
// Ylitie, Karras, Laine, "Efficient Incoherent Ray Traversal on GPUs Through
// Compressed Wide BVHs" (HPG 2017).
//
// Node layout (all little endian):
//   float    p[3]        origin of the local grid (slightly below the node's min corner)
//   uint8_t  e[3]        per-axis grid scale exponent: cell size = 2^(e-127)
//   uint8_t  imask       bit s set <=> slot s holds an inner child
//   uint32_t child_base  index of the first inner child (inner children are contiguous, slot order)
//   uint32_t tri_base    index of the first triangle of this node's leaf children
//   uint8_t  meta[8]     per slot: 0 empty | inner: 001sssss (sssss = 24+slot)
//                        | leaf: unary triangle count (001/011/111) in the top 3 bits, first
//                          triangle (relative to tri_base) in the low 5 bits
//   uint8_t  qlo[3][8], qhi[3][8]   child boxes in grid cells relative to p
// Child boxes decode to supersets of the true boxes with at least one grid
// cell of slack on every side, so the quantised slab test stays conservative
// under float rounding (the hit decision itself is made by the exact
// watertight triangle test, never by a box).
//
// Triangles are stored in leaf order as three float4: (p0, prim id) (p1,
// material id | flip << 16) (p2, light id): everything shading needs sits in
// the 48 bytes traversal already touched.
#ifndef B200PT_BVH8_H
#define B200PT_BVH8_H

#include <cstdint>
#include <vector>

namespace b200pt {

struct alignas(16) Bvh8Node {
    float p[3];
    uint8_t e[3];
    uint8_t imask;
    uint32_t child_base;
    uint32_t tri_base;
    uint8_t meta[8];
    uint8_t qlo[3][8];
    uint8_t qhi[3][8];
};
static_assert(sizeof(Bvh8Node) == 80, "Bvh8Node must be 80 bytes");

struct alignas(16) TriRecord {  // 48 bytes
    float p0[3];
    uint32_t prim;
    float p1[3];
    uint32_t mat_flags;  // material id | (flip_normal << 16) | (degenerate << 17)
    float p2[3];
    int32_t light;
};
static_assert(sizeof(TriRecord) == 48, "TriRecord must be 48 bytes");

struct Bvh8 {
    std::vector<Bvh8Node> nodes;     // node 0 is the root
    std::vector<TriRecord> tris;     // leaf order; degenerate triangles (never hittable) at the end
    std::vector<uint32_t> prim_to_tri;  // original triangle index -> position in `tris`
    uint32_t n_in_leaves = 0;        // triangles referenced by leaves
    int max_depth = 0;               // depth of the wide tree (root = 1)
};

// Builds the wide BVH on the host: binned-SAH binary build (multi-threaded),
// greedy collapse to 8-wide, octant-ordered slot assignment, quantisation.
// `degenerate[i]` marks triangles the reference can never hit
// (shapes/triangle.cpp:304-312); they are kept out of the leaves.
void build_bvh8(const float *vertices, int64_t n_tris, const int32_t *material_id, const int32_t *light_id,
                const uint8_t *flip, const uint8_t *degenerate, int n_threads, Bvh8 *out);

// Structural self-check used by the library after every build: every leaf
// triangle's exact bounds lie inside the decoded box of its slot and every
// inner child's decoded box lies inside its parent's slot box.  Returns the
// number of violations.
int64_t validate_bvh8(const Bvh8 &bvh);

}  // namespace b200pt
#endif
