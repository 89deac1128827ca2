// wbvh_gpu.cu -- on-device build of the 7-wide BVH (lbvh.cuh has the per-element
// steps and the rationale).  Selected with b200pt_ctx_set_option(ctx,
// "gpu_bvh_build", 1) or B200PT_BVH_BUILD=gpu; the default stays the host SAH
// builder (wbvh_build.cpp), whose trees traverse faster.  The only library call
// is the radix sort of the (Morton key, triangle) pairs (CUB); everything else
// is kernels over lbvh.cuh.
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <vector>

#include "wbvh_gpu.h"
#include "lbvh.cuh"

namespace b200pt {
namespace {

__global__ void k_lbvh_prep(const LbvhCtx c) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.n) lbvh_prep(c, i);
}
__global__ void k_lbvh_key(const LbvhCtx c) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < c.m) lbvh_key(c, k);
}
__global__ void k_lbvh_karras(const LbvhCtx c) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.m - 1) lbvh_karras(c, i);
}
__global__ void k_lbvh_fit(const LbvhCtx c) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < c.m) lbvh_fit_from_leaf(c, j);
}
__global__ void k_lbvh_collapse(const LbvhCtx c, uint32_t n_items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_items) lbvh_collapse(c, c.q_in[i]);
}
__global__ void k_lbvh_leftover(const LbvhCtx c) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c.n) lbvh_leftover(c, i);
}

struct Scratch {
    std::vector<void *> ptrs;
    cudaError_t err = cudaSuccess;
    template <typename T>
    T *alloc(size_t count) {
        void *p = nullptr;
        if (err == cudaSuccess) err = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (p) ptrs.push_back(p);
        return static_cast<T *>(p);
    }
    ~Scratch() {
        for (void *p : ptrs) cudaFree(p);
    }
};

inline unsigned blocks(int64_t n, int per) { return (unsigned)std::max<int64_t>(1, (n + per - 1) / per); }

}  // namespace

#define GB_TRY(call)                                  \
    do {                                              \
        cudaError_t e_ = (call);                      \
        if (e_ != cudaSuccess) {                      \
            snprintf(err, err_len, "%s: %s", #call, cudaGetErrorString(e_)); \
            return false;                             \
        }                                             \
    } while (0)

bool build_wbvh_gpu(const GpuBuildInput &in, cudaStream_t st, GpuBuildOutput *out, char *err, size_t err_len) {
    const int64_t n = in.n_tris;
    Scratch tmp;
    LbvhCtx c;
    memset(&c, 0, sizeof(c));
    c.n = n;
    // ---- upload the descriptor arrays
    float *d_vertices = tmp.alloc<float>((size_t)n * 9);
    int32_t *d_mat = tmp.alloc<int32_t>((size_t)n), *d_light = in.light_id ? tmp.alloc<int32_t>((size_t)n) : nullptr;
    uint8_t *d_flip = in.flip ? tmp.alloc<uint8_t>((size_t)n) : nullptr;
    uint8_t *d_vf = in.vertex_flags ? tmp.alloc<uint8_t>((size_t)n) : nullptr;
    float *d_uvs = in.uvs ? tmp.alloc<float>((size_t)n * 6) : nullptr;
    c.valid_idx = tmp.alloc<uint32_t>((size_t)n);
    uint32_t *d_counters = tmp.alloc<uint32_t>(8);  // n_valid, n_nodes, n_tris, q_out_count
    c.cbounds = tmp.alloc<int32_t>(12);
    if (tmp.err != cudaSuccess) GB_TRY(tmp.err);
    if (n) {
        GB_TRY(cudaMemcpyAsync(d_vertices, in.vertices, (size_t)n * 36, cudaMemcpyHostToDevice, st));
        GB_TRY(cudaMemcpyAsync(d_mat, in.material_id, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        if (d_light) GB_TRY(cudaMemcpyAsync(d_light, in.light_id, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        if (d_flip) GB_TRY(cudaMemcpyAsync(d_flip, in.flip, (size_t)n, cudaMemcpyHostToDevice, st));
        if (d_vf) GB_TRY(cudaMemcpyAsync(d_vf, in.vertex_flags, (size_t)n, cudaMemcpyHostToDevice, st));
        if (d_uvs) GB_TRY(cudaMemcpyAsync(d_uvs, in.uvs, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    }
    c.vertices = d_vertices;
    c.material_id = d_mat;
    c.light_id = d_light;
    c.flip = d_flip;
    c.vertex_flags = d_vf;
    c.uvs = d_uvs;
    c.has_normals = in.has_normals;
    c.has_uvs = in.uvs != nullptr;
    c.n_valid = d_counters + 0;
    c.n_nodes = d_counters + 1;
    c.n_tris = d_counters + 2;
    c.q_out_count = d_counters + 3;
    const uint32_t counters0[8] = {0, 1, 0, 0, 0, 0, 0, 0};  // node 0 = root
    const int32_t bounds0[12] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    GB_TRY(cudaMemcpyAsync(d_counters, counters0, sizeof(counters0), cudaMemcpyHostToDevice, st));
    GB_TRY(cudaMemcpyAsync(c.cbounds, bounds0, sizeof(bounds0), cudaMemcpyHostToDevice, st));
    // outputs (final size of the triangle array is known: every triangle gets a record)
    GB_TRY(cudaMalloc(&out->d_tris, std::max<size_t>(1, (size_t)n) * sizeof(TriRecord)));
    GB_TRY(cudaMalloc(&out->d_prim_to_tri, std::max<size_t>(1, (size_t)n) * sizeof(uint32_t)));
    c.tris = static_cast<TriRecord *>(out->d_tris);
    c.prim_to_tri = static_cast<uint32_t *>(out->d_prim_to_tri);

    // ---- 1. filter + centroid bounds
    if (n) k_lbvh_prep<<<blocks(n, 256), 256, 0, st>>>(c);
    uint32_t m32 = 0;
    GB_TRY(cudaMemcpyAsync(&m32, c.n_valid, 4, cudaMemcpyDeviceToHost, st));
    GB_TRY(cudaStreamSynchronize(st));
    const int64_t m = m32;
    c.m = m;
    {   // bounds of the triangles in the tree -> smallest cell of the quantisation grids (wbvh.h)
        int32_t tb[6];
        GB_TRY(cudaMemcpy(tb, c.cbounds + 6, sizeof(tb), cudaMemcpyDeviceToHost));
        float absmax = 0.f;
        for (int a = 0; a < 3; ++a) {
            out->bounds_lo[a] = m ? lb_ordered_to_float(tb[a]) : 0.f;
            out->bounds_hi[a] = m ? lb_ordered_to_float(tb[3 + a]) : 0.f;
            absmax = std::max(absmax, std::max(std::fabs(out->bounds_lo[a]), std::fabs(out->bounds_hi[a])));
        }
        c.cell_floor = B200PT_CELL_FLOOR * absmax;
    }
    // wide nodes: every node other than the root holds at least two triangles and a node with a large inner child
    // always has 7 children, hence fewer than m nodes (lbvh.cuh)
    const size_t node_cap = (size_t)m + 16;
    WbvhNode *d_nodes_tmp = tmp.alloc<WbvhNode>(node_cap);
    uint32_t *d_tri_base_tmp = tmp.alloc<uint32_t>(node_cap);
    c.nodes = d_nodes_tmp;
    c.tri_base = d_tri_base_tmp;
    c.node_cap = node_cap;
    int depth = 1;
    if (m > 0) {
        c.keys = tmp.alloc<uint64_t>((size_t)m);
        c.sorted = tmp.alloc<uint32_t>((size_t)m);
        uint64_t *keys_alt = tmp.alloc<uint64_t>((size_t)m);
        uint32_t *sorted_alt = tmp.alloc<uint32_t>((size_t)m);
        c.left = tmp.alloc<int32_t>((size_t)m);
        c.right = tmp.alloc<int32_t>((size_t)m);
        c.parent = tmp.alloc<int32_t>((size_t)2 * m);
        c.first = tmp.alloc<int32_t>((size_t)m);
        c.last = tmp.alloc<int32_t>((size_t)m);
        c.nbox = tmp.alloc<float>((size_t)m * 6);
        c.arrivals = tmp.alloc<uint32_t>((size_t)m);
        LbvhItem *q[2] = {tmp.alloc<LbvhItem>((size_t)m + 1), tmp.alloc<LbvhItem>((size_t)m + 1)};
        if (tmp.err != cudaSuccess) GB_TRY(tmp.err);
        // ---- 2. Morton keys, sorted (CUB double buffer: results end up in c.keys / c.sorted after the swap below)
        k_lbvh_key<<<blocks(m, 256), 256, 0, st>>>(c);
        cub::DoubleBuffer<uint64_t> kb(c.keys, keys_alt);
        cub::DoubleBuffer<uint32_t> vb(c.sorted, sorted_alt);
        size_t sort_bytes = 0;
        GB_TRY(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, kb, vb, (int)m, 0, 63, st));
        void *sort_tmp = tmp.alloc<uint8_t>(sort_bytes);
        if (tmp.err != cudaSuccess) GB_TRY(tmp.err);
        GB_TRY(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, kb, vb, (int)m, 0, 63, st));
        c.keys = kb.Current();
        c.sorted = vb.Current();
        // ---- 3./4. binary radix tree + bounds
        if (m > 1) {
            k_lbvh_karras<<<blocks(m - 1, 256), 256, 0, st>>>(c);
            k_lbvh_fit<<<blocks(m, 256), 256, 0, st>>>(c);
        }
        // ---- 5. collapse, one launch per level of the wide tree
        LbvhItem root;
        root.node2 = m > 1 ? 0 : ~0;
        root.wide = 0;
        GB_TRY(cudaMemcpyAsync(q[0], &root, sizeof(root), cudaMemcpyHostToDevice, st));
        uint32_t n_items = 1;
        int cur = 0;
        while (n_items) {
            GB_TRY(cudaMemsetAsync(c.q_out_count, 0, 4, st));
            c.q_in = q[cur];
            c.q_out = q[cur ^ 1];
            k_lbvh_collapse<<<blocks(n_items, 128), 128, 0, st>>>(c, n_items);
            GB_TRY(cudaMemcpyAsync(&n_items, c.q_out_count, 4, cudaMemcpyDeviceToHost, st));
            GB_TRY(cudaStreamSynchronize(st));
            cur ^= 1;
            if (n_items) ++depth;
            if (depth > 4096) {
                snprintf(err, err_len, "gpu bvh build: runaway tree depth");
                return false;
            }
        }
    } else {
        // empty scene: a root with no children (same record as the host builder)
        WbvhNode node;
        memset(&node, 0, sizeof(node));
        const int none[B200PT_WIDTH] = {-1, -1, -1, -1, -1, -1, -1};
        wbvh_encode_node(nullptr, none, nullptr, 1.f, &node);
        GB_TRY(cudaMemcpyAsync(d_nodes_tmp, &node, sizeof(node), cudaMemcpyHostToDevice, st));
        GB_TRY(cudaMemsetAsync(d_tri_base_tmp, 0, 4, st));
        GB_TRY(cudaStreamSynchronize(st));
    }
    uint32_t counts[3];
    GB_TRY(cudaMemcpyAsync(counts, d_counters, sizeof(counts), cudaMemcpyDeviceToHost, st));
    GB_TRY(cudaStreamSynchronize(st));
    out->n_in_leaves = counts[2];
    // ---- 6. records for the triangles that are not in the tree
    if (n) k_lbvh_leftover<<<blocks(n, 256), 256, 0, st>>>(c);
    out->n_nodes = counts[1];
    out->n_tris = (uint64_t)n;
    out->max_depth = depth;
    if ((size_t)out->n_nodes > node_cap) {
        snprintf(err, err_len, "gpu bvh build: node estimate exceeded (%u > %zu)", counts[1], node_cap);
        return false;
    }
    GB_TRY(cudaMalloc(&out->d_nodes, std::max<size_t>(1, out->n_nodes) * sizeof(WbvhNode)));
    GB_TRY(cudaMalloc(&out->d_tri_base, std::max<size_t>(1, out->n_nodes) * sizeof(uint32_t)));
    GB_TRY(cudaMemcpyAsync(out->d_nodes, d_nodes_tmp, out->n_nodes * sizeof(WbvhNode), cudaMemcpyDeviceToDevice, st));
    GB_TRY(cudaMemcpyAsync(out->d_tri_base, d_tri_base_tmp, out->n_nodes * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    GB_TRY(cudaStreamSynchronize(st));
    GB_TRY(cudaGetLastError());
    return true;
}

}  // namespace b200pt
