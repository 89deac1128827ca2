// tests/emu/cuda_runtime.h -- TEST INFRASTRUCTURE, never part of libb200pt.so.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile the product's api.cu and kernels.cu as plain C++
// (tests/emu/Makefile, -DB200PT_HOST_EMU): "device" memory is host memory, a kernel launch runs the kernel
// function once per thread index (blocks concurrently on a few host threads, the threads of a block one after the
// other; atomics are real), and the warp-aggregated helpers of kernels.cu take their one-lane form.  The resulting libb200pt_hostcheck.so exports the C ABI of include/b200pt.h so that the
// parity tests can pre-flight the host logic (scene / render set-up, wavefront sequencing) and the scalar logic
// of every kernel against the oracle on a box without a GPU, the way tests/host_preflight.cpp pre-flights the
// math headers.  It is a checker of the sources, not a renderer: the package never loads it, nothing ships it,
// and it is ~1000x slower than the reference it is checked against.  The warp-synchronous traversal kernel
// k_trace is the one piece it cannot execute; its rays go through traverse_wbvh (wbvh_traverse.cuh), the same
// per-ray routine the kernel's lanes step through.
#ifndef B200PT_EMU_CUDA_RUNTIME_H
#define B200PT_EMU_CUDA_RUNTIME_H
#ifndef B200PT_HOST_EMU
#error "tests/emu/cuda_runtime.h is only for the -DB200PT_HOST_EMU check build"
#endif
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __constant__ static
#define __shared__ static

struct float2 {
    float x, y;
};
struct float4 {
    float x, y, z, w;
};
struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace b200pt_emu {
inline thread_local uint3 thread_idx = {0, 0, 0}, block_idx = {0, 0, 0};
inline thread_local dim3 block_dim, grid_dim;
inline int worker_count() {
    static const int n = []() {
        const char *e = getenv("B200PT_EMU_THREADS");
        const int hw = (int)std::thread::hardware_concurrency();
        return std::max(1, e ? atoi(e) : std::min(hw > 0 ? hw : 1, 16));
    }();
    return n;
}
// One kernel launch: the blocks are dealt to a few host threads (like CTAs to SMs: any order, concurrently -- the
// kernels' atomics are real atomics here), the threads of a block run one after the other.
template <class F>
inline void launch(dim3 grid, dim3 block, F &&body) {
    const unsigned long long n_blocks = (unsigned long long)grid.x * grid.y * grid.z;
    std::atomic<unsigned long long> next{0};
    auto worker = [&]() {
        grid_dim = grid;
        block_dim = block;
        for (;;) {
            const unsigned long long b = next.fetch_add(1);
            if (b >= n_blocks) break;
            block_idx = uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long long)grid.x * grid.y))};
            for (unsigned tz = 0; tz < block.z; ++tz)
                for (unsigned ty = 0; ty < block.y; ++ty)
                    for (unsigned tx = 0; tx < block.x; ++tx) {
                        thread_idx = uint3{tx, ty, tz};
                        body();
                    }
        }
    };
    const int n_workers = (int)std::min<unsigned long long>((unsigned long long)worker_count(), n_blocks);
    if (n_workers <= 1) {
        worker();
        return;
    }
    std::vector<std::thread> pool;
    for (int i = 1; i < n_workers; ++i) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}
}  // namespace b200pt_emu
#define threadIdx (::b200pt_emu::thread_idx)
#define blockIdx (::b200pt_emu::block_idx)
#define blockDim (::b200pt_emu::block_dim)
#define gridDim (::b200pt_emu::grid_dim)

inline uint32_t __float_as_uint(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
inline float __uint_as_float(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline float atomicAdd(float *p, float v) {
    uint32_t *u = reinterpret_cast<uint32_t *>(p);
    uint32_t old = __atomic_load_n(u, __ATOMIC_RELAXED);
    for (;;) {
        float f;
        memcpy(&f, &old, 4);
        f += v;
        uint32_t want;
        memcpy(&want, &f, 4);
        if (__atomic_compare_exchange_n(u, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) break;
    }
    float r;
    memcpy(&r, &old, 4);
    return r;
}
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }

// ---- runtime API: one "device" whose memory is the host's
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
};
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated allocation failure"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) {
    *n = 1;
    return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "host check build (no GPU)");
    p->major = 10;
    p->multiProcessorCount = 8;  // the persistent kernels size their grids from this: enough blocks for the host threads
    return cudaSuccess;
}
// B200PT_EMU_POISON=<byte>: fresh "device" memory is filled with that byte (0xff: NaNs / huge indices), so a kernel
// that reads memory nothing wrote shows up as a changed image or a crash instead of passing by luck
template <class T>
inline cudaError_t cudaMalloc(T **p, size_t n) {
    *p = static_cast<T *>(malloc(n ? n : 1));
    if (!*p) return cudaErrorMemoryAllocation;
    static const char *poison = getenv("B200PT_EMU_POISON");
    if (poison) memset(*p, (int)strtol(poison, nullptr, 0), n ? n : 1);
    return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMallocHost(T **p, size_t n) {
    return cudaMalloc(p, n);
}
inline cudaError_t cudaFree(void *p) {
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void *p) {
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) {
    if (n) memset(d, v, n);
    return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMemcpyToSymbolAsync(T &symbol, const void *s, size_t n, size_t off, cudaMemcpyKind, cudaStream_t) {
    memcpy(reinterpret_cast<char *>(&symbol) + off, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    *s = nullptr;
    return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
    *e = nullptr;
    return cudaSuccess;
}
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) {
    *ms = 0.f;
    return cudaSuccess;
}
#endif
