// wbvh_gpu.h -- interface of the on-device BVH builder (wbvh_gpu.cu, lbvh.cuh).
#ifndef B200PT_WBVH_GPU_H
#define B200PT_WBVH_GPU_H

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace b200pt {

struct GpuBuildInput {  // host pointers, the arrays of b200pt_scene_desc
    const float *vertices;
    int64_t n_tris;
    const int32_t *material_id, *light_id;
    const uint8_t *flip, *vertex_flags;
    const float *uvs;
    int has_normals;
};

struct GpuBuildOutput {  // device allocations (cudaMalloc) owned by the caller afterwards, also on failure
    void *d_nodes = nullptr;        // WbvhNode[n_nodes]
    void *d_tri_base = nullptr;     // uint32_t[n_nodes]
    void *d_tris = nullptr;         // TriRecord[n_tris], leaf order, never-hittable triangles last
    void *d_prim_to_tri = nullptr;  // uint32_t[n_tris]
    uint64_t n_nodes = 0, n_tris = 0;
    uint32_t n_in_leaves = 0;
    int max_depth = 0;
    float bounds_lo[3] = {0, 0, 0}, bounds_hi[3] = {0, 0, 0};  // bounds of the triangles in the tree
};

// Builds the tree on the device of the current context, on `st`; returns false and a message in `err` on failure.
bool build_wbvh_gpu(const GpuBuildInput &in, cudaStream_t st, GpuBuildOutput *out, char *err, size_t err_len);

}  // namespace b200pt
#endif
