// pr_compat.h -- one source, two compilations.
//   nvcc (product):     real CUDA for sm_100a. This is the ONLY thing libplayrho_b200.so contains.
//   g++ -DPR_HOST_EMUL: a serial "one thread after another" execution of the same kernels, built
//                       ONLY by tests/ (tests/_emul/) to check host logic and kernel bodies on the
//                       GPU-less CI box. It is never built into, loaded by, or reachable from the
//                       product library or the python package (no CPU fallback exists).
#ifndef PR_COMPAT_H
#define PR_COMPAT_H

#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cmath>

#ifndef PR_HOST_EMUL
// ------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#define PR_HD __host__ __device__ __forceinline__
#define PR_D __device__ __forceinline__
#define PR_KERNEL __global__ void
#define PR_KERNEL_LB(threads, blocks) __global__ void __launch_bounds__(threads, blocks)
#define PR_TID (blockIdx.x * blockDim.x + threadIdx.x)
#define PR_NTHREADS (gridDim.x * blockDim.x)
typedef cudaStream_t pr_stream_t;
typedef cudaEvent_t pr_event_t;

#define PR_BLOCK 256
// warp-cooperative loops (a "warp" of the host emulation is one thread)
#define PR_WARP 32u
#define PR_LANE (threadIdx.x & 31u)
#define PR_WARP_ID (PR_TID >> 5)
#define PR_NWARPS (PR_NTHREADS >> 5)
#define PR_WARP_ALL(p) (__all_sync(0xffffffffu, (p)) != 0)
#define PR_WARP_ANY(p) (__any_sync(0xffffffffu, (p)) != 0)
#define PR_BALLOT(p) __ballot_sync(0xffffffffu, (p))
#define PR_SHFL(v, src) __shfl_sync(0xffffffffu, (v), (src))
#define PR_SYNCWARP() __syncwarp()
#define PR_LANEMASK_LT ((1u << (threadIdx.x & 31u)) - 1u)
#define PR_WARP_SUM(v) __reduce_add_sync(0xffffffffu, (v))
#define PR_DN __device__ __noinline__
#define PR_GRID_CONST const __grid_constant__ // (a kernel parameter whose address may be taken without a local copy)
// Every kernel launch is a programmatic dependent launch (sm_90+, below): the grid may be SCHEDULED while its predecessor in the
// stream is still draining, and every kernel body starts with PR_PDL_ENTER() -- let my own dependents be scheduled, then wait
// for my predecessor to complete (and its writes to be visible) -- before it touches global memory. A step is ~300 dependent
// launches; this hides the launch gap of each.
#define PR_LAUNCH(kernel, n, strm_, ...)                                                           \
    do {                                                                                           \
        const unsigned pr_n_ = static_cast<unsigned>(n);                                           \
        if (pr_n_ > 0) PR_LAUNCH_GRID(kernel, (pr_n_ + PR_BLOCK - 1) / PR_BLOCK, PR_BLOCK, strm_, __VA_ARGS__); \
    } while (0)
#define PR_LAUNCH_GRID(kernel, grid, block, strm_, ...)                                            \
    do {                                                                                           \
        cudaLaunchConfig_t pr_cfg_ = {};                                                           \
        pr_cfg_.gridDim = dim3(grid); pr_cfg_.blockDim = dim3(block); pr_cfg_.stream = strm_;      \
        cudaLaunchAttribute pr_at_[1];                                                             \
        pr_at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                         \
        pr_at_[0].val.programmaticStreamSerializationAllowed = 1;                                  \
        pr_cfg_.attrs = pr_at_; pr_cfg_.numAttrs = 1;                                              \
        cudaLaunchKernelEx(&pr_cfg_, kernel, __VA_ARGS__);                                         \
        pr::count_launch();                                                                        \
    } while (0)
#define PR_LAUNCH_GRID_SMEM(kernel, grid, block, smemBytes, strm_, ...)                            \
    do {                                                                                           \
        cudaLaunchConfig_t pr_cfg_ = {};                                                           \
        pr_cfg_.gridDim = dim3(grid); pr_cfg_.blockDim = dim3(block); pr_cfg_.stream = strm_;      \
        pr_cfg_.dynamicSmemBytes = smemBytes;                                                      \
        cudaLaunchAttribute pr_at_[1];                                                             \
        pr_at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                         \
        pr_at_[0].val.programmaticStreamSerializationAllowed = 1;                                  \
        pr_cfg_.attrs = pr_at_; pr_cfg_.numAttrs = 1;                                              \
        cudaLaunchKernelEx(&pr_cfg_, kernel, __VA_ARGS__);                                         \
        pr::count_launch();                                                                        \
    } while (0)
// (the per-colour solver kernels go further: they fetch their immutable row planes BEFORE the wait)
#define PR_LAUNCH_PDL PR_LAUNCH
__device__ __forceinline__ void pr_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pr_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define PR_PDL_ENTER() do { pr_pdl_launch_dependents(); pr_pdl_wait(); } while (0)

#else
// ------------------------------------------------------------------ host emulation (tests only)
#include <algorithm>
#include <cstdlib>
#define PR_HD inline
#define PR_D inline
#define PR_KERNEL void
#define PR_KERNEL_LB(threads, blocks) void
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
namespace pr { extern thread_local unsigned emul_tid, emul_nthreads; }
#define PR_TID (pr::emul_tid)
#define PR_NTHREADS (pr::emul_nthreads)
#define PR_WARP 1u
#define PR_LANE 0u
#define PR_WARP_ID (pr::emul_tid)
#define PR_NWARPS (pr::emul_nthreads)
#define PR_WARP_ALL(p) (p)
#define PR_WARP_ANY(p) (p)
#define PR_BALLOT(p) ((p) ? 1u : 0u)
#define PR_SHFL(v, src) (v)
#define PR_SYNCWARP() do { } while (0)
#define PR_LANEMASK_LT 0u
#define PR_WARP_SUM(v) (v)
#define PR_DN inline
#define PR_GRID_CONST const
typedef int pr_stream_t;
typedef int pr_event_t;
#define PR_LAUNCH(kernel, n, stream, ...)                                                          \
    do {                                                                                           \
        const unsigned pr_n_ = static_cast<unsigned>(n);                                           \
        pr::emul_nthreads = pr_n_;                                                                 \
        for (unsigned pr_t_ = 0; pr_t_ < pr_n_; ++pr_t_) { pr::emul_tid = pr_t_; kernel(__VA_ARGS__); } \
        if (pr_n_ > 0) pr::count_launch();                                                         \
    } while (0)
#define PR_LAUNCH_PDL PR_LAUNCH
inline void pr_pdl_wait() {}
inline void pr_pdl_launch_dependents() {}
#define PR_PDL_ENTER() do { } while (0)
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline int __clzll(unsigned long long v) { return v ? __builtin_clzll(v) : 64; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(unsigned long long v) { return __builtin_ffsll(static_cast<long long>(v)); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __syncthreads() {}
#endif

namespace pr {
// host+device bit casts
PR_HD unsigned f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    unsigned u; std::memcpy(&u, &f, 4); return u;
#endif
}
PR_HD float u2f(unsigned u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; std::memcpy(&f, &u, 4); return f;
#endif
}
void count_launch();
} // namespace pr

#endif // PR_COMPAT_H
