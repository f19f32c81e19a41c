// pr_prims.cuh -- hand-written device-wide primitives used by the step pipeline:
//   scan_exclusive_u32 : 3-kernel tiled exclusive prefix sum (warp-shuffle block scan)
//   radix_sort_pairs   : stable LSD radix sort, 8 bits per pass (block histogram -> scan -> stable
//                        in-block ranking with __match_any_sync -> scatter); 32- or 64-bit keys
// No CUB/Thrust. In the tests-only host emulation (PR_HOST_EMUL) these two are replaced by their
// std:: equivalents, since the emulation runs threads one after another and cannot express
// __syncthreads()/warp intrinsics.
#ifndef PR_PRIMS_CUH
#define PR_PRIMS_CUH
#include "pr_compat.h"

namespace pr {

#ifndef PR_HOST_EMUL

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned warp_incl_scan(unsigned v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across a block of SCAN_THREADS; returns exclusive prefix,
// *total = block total. smem: 32 uints.
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned* smem, unsigned* total)
{
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned incl = warp_incl_scan(v);
    if (lane == 31) smem[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const unsigned nw = blockDim.x >> 5;
        unsigned w = (lane < nw) ? smem[lane] : 0u;
        const unsigned wi = warp_incl_scan(w);
        smem[lane] = wi - w; // exclusive warp offsets
        if (lane == nw - 1) smem[32] = wi;
    }
    __syncthreads();
    const unsigned r = smem[wid] + incl - v;
    *total = smem[32];
    __syncthreads();
    return r;
}

__global__ void k_scan_tile_sums(const unsigned* __restrict__ in, unsigned* __restrict__ tileSums, unsigned n)
{
    PR_PDL_ENTER();
    __shared__ unsigned smem[33];
    const unsigned base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { const unsigned j = base + i; if (j < n) s += in[j]; }
    unsigned total;
    block_excl_scan(s, smem, &total);
    if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

// single block: exclusive scan of tileSums in place (sequential over chunks with a carry)
__global__ void k_scan_tile_offsets(unsigned* __restrict__ tileSums, unsigned nTiles, unsigned* __restrict__ totalOut)
{
    PR_PDL_ENTER();
    __shared__ unsigned smem[33];
    unsigned carry = 0;
    for (unsigned base = 0; base < nTiles; base += blockDim.x) {
        const unsigned j = base + threadIdx.x;
        const unsigned v = (j < nTiles) ? tileSums[j] : 0u;
        unsigned total;
        const unsigned ex = block_excl_scan(v, smem, &total);
        if (j < nTiles) tileSums[j] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && totalOut) *totalOut = carry;
}

__global__ void k_scan_apply(const unsigned* __restrict__ in, unsigned* __restrict__ out,
                             const unsigned* __restrict__ tileOffsets, unsigned n)
{
    PR_PDL_ENTER();
    __shared__ unsigned smem[33];
    const unsigned base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS];
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { const unsigned j = base + i; v[i] = (j < n) ? in[j] : 0u; s += v[i]; }
    unsigned total;
    unsigned ex = block_excl_scan(s, smem, &total) + tileOffsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { const unsigned j = base + i; if (j < n) out[j] = ex; ex += v[i]; }
}

// small inputs: ONE block walks the array in chunks of blockDim.x * SCAN_ITEMS with a running carry (one launch
// instead of three; the radix passes over a few ten thousand candidate pairs scan 256 x ~100 counters each)
constexpr unsigned SCAN_SMALL_MAX = 8192u; // one chunk of the block (1024 threads x 8): beyond it the three tiled kernels win (31 vs 16 us at 25k, ncu)
__global__ void k_scan_small(const unsigned* in, unsigned* out, unsigned n, unsigned* totalOut) // (in may alias out)
{
    PR_PDL_ENTER();
    __shared__ unsigned smem[33];
    unsigned carry = 0;
    for (unsigned chunk = 0; chunk < n; chunk += blockDim.x * SCAN_ITEMS) {
        const unsigned base = chunk + threadIdx.x * SCAN_ITEMS;
        unsigned v[SCAN_ITEMS];
        unsigned sum = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) { const unsigned j = base + i; v[i] = (j < n) ? in[j] : 0u; sum += v[i]; }
        unsigned total;
        unsigned ex = block_excl_scan(sum, smem, &total) + carry;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) { const unsigned j = base + i; if (j < n) out[j] = ex; ex += v[i]; }
        carry += total;
    }
    if (threadIdx.x == 0 && totalOut) *totalOut = carry;
}

// scratch must hold ceil(n / SCAN_TILE) uints. in may alias out. totalOut (device) may be null.
inline void scan_exclusive_u32(const unsigned* in, unsigned* out, unsigned n, unsigned* totalOut,
                               unsigned* scratch, cudaStream_t stream)
{
    if (n == 0) {
        if (totalOut) { cudaMemsetAsync(totalOut, 0, sizeof(unsigned), stream); count_launch(); } // (a device counter changed: see read_counters)
        return;
    }
    if (n <= SCAN_SMALL_MAX) { PR_LAUNCH_GRID(k_scan_small, 1, 1024, stream, in, out, n, totalOut); return; }
    const unsigned nTiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    PR_LAUNCH_GRID(k_scan_tile_sums, nTiles, SCAN_THREADS, stream, in, scratch, n);
    PR_LAUNCH_GRID(k_scan_tile_offsets, 1, SCAN_THREADS, stream, scratch, nTiles, totalOut);
    PR_LAUNCH_GRID(k_scan_apply, nTiles, SCAN_THREADS, stream, in, out, scratch, n);
}
inline size_t scan_scratch_elems(unsigned n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

// ---------------------------------------------------------------- radix sort
constexpr int RS_THREADS = 256;
constexpr int RS_CHUNKS = 16;                       // chunks of RS_THREADS keys per block
constexpr int RS_TILE = RS_THREADS * RS_CHUNKS;     // 4096 keys per block

// CHUNKS: chunks of RS_THREADS keys per block -- RS_CHUNKS for large inputs, RS_CHUNKS_SMALL for mid-size ones (a few ten
// thousand keys on 4096-key tiles is a dozen blocks, each ranking 16 chunks one after another: 22 us per pass measured)
constexpr int RS_CHUNKS_SMALL = 2;
constexpr unsigned RS_SMALL_MAX = 262144u;
template <class K, int CHUNKS>
__global__ void k_rs_hist(const K* __restrict__ keys, unsigned n, int shift, unsigned* __restrict__ hist, unsigned nBlocks)
{
    PR_PDL_ENTER();
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned base = blockIdx.x * (CHUNKS * RS_THREADS);
#pragma unroll 4
    for (int c = 0; c < CHUNKS; ++c) {
        const unsigned j = base + c * RS_THREADS + threadIdx.x;
        if (j < n) atomicAdd(&h[static_cast<unsigned>(keys[j] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nBlocks + blockIdx.x] = h[threadIdx.x]; // digit-major so one scan gives offsets
}

template <class K, int CHUNKS>
__global__ void k_rs_scatter(const K* __restrict__ keysIn, const unsigned* __restrict__ valsIn,
                             K* __restrict__ keysOut, unsigned* __restrict__ valsOut, unsigned n, int shift,
                             const unsigned* __restrict__ offsets, unsigned nBlocks)
{
    PR_PDL_ENTER();
    __shared__ unsigned running[256];                 // next output slot per digit for this block
    __shared__ unsigned warpCnt[RS_THREADS / 32][256];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    running[threadIdx.x] = offsets[threadIdx.x * nBlocks + blockIdx.x];
    const unsigned base = blockIdx.x * (CHUNKS * RS_THREADS);
    for (int c = 0; c < CHUNKS; ++c) {
        if (base + c * RS_THREADS >= n) break;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) warpCnt[w][threadIdx.x] = 0;
        __syncthreads();
        const unsigned j = base + c * RS_THREADS + threadIdx.x;
        const bool valid = j < n;
        K key = 0; unsigned val = 0; unsigned d = 0xFFFFFFFFu;
        if (valid) { key = keysIn[j]; val = valsIn ? valsIn[j] : j; d = static_cast<unsigned>(key >> shift) & 0xFFu; }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned rankInWarp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rankInWarp == 0) warpCnt[wid][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            unsigned pos = running[d] + rankInWarp;
            for (unsigned w = 0; w < wid; ++w) pos += warpCnt[w][d];
            keysOut[pos] = key;
            if (valsOut) valsOut[pos] = val;
        }
        __syncthreads();
        unsigned add = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) add += warpCnt[w][threadIdx.x];
        running[threadIdx.x] += add;
        __syncthreads();
    }
}

struct RadixScratch {
    unsigned* hist = nullptr;     // 256 * nBlocks
    unsigned* scan = nullptr;     // scan scratch
    size_t histCap = 0, scanCap = 0;
};
inline unsigned rs_blocks(unsigned n) { return (n + RS_TILE - 1) / RS_TILE; }

// Sort (keys, vals) by bits [beginBit, endBit) of the key; ping-pongs between the A and B buffers.
// valsA == nullptr on input means "values are the element indices" for the first pass.
// Returns true if the result ended up in the B buffers.
template <class K>
inline bool radix_sort_pairs(K* keysA, unsigned* valsA, K* keysB, unsigned* valsB, unsigned n, int beginBit,
                             int endBit, bool valsAreIota, RadixScratch& s, cudaStream_t stream)
{
    if (n == 0) return false;
    const unsigned smallBlocks = (n + RS_CHUNKS_SMALL * RS_THREADS - 1) / (RS_CHUNKS_SMALL * RS_THREADS);
    const bool small = n <= RS_SMALL_MAX && 256ull * smallBlocks <= s.histCap;
    const unsigned nBlocks = small ? smallBlocks : rs_blocks(n);
    bool inA = true;
    bool iota = valsAreIota;
    for (int shift = beginBit; shift < endBit; shift += 8) {
        K* kin = inA ? keysA : keysB;
        K* kout = inA ? keysB : keysA;
        unsigned* vin = inA ? valsA : valsB;
        unsigned* vout = inA ? valsB : valsA;
        if (small) PR_LAUNCH_GRID((k_rs_hist<K, RS_CHUNKS_SMALL>), nBlocks, RS_THREADS, stream, kin, n, shift, s.hist, nBlocks);
        else PR_LAUNCH_GRID((k_rs_hist<K, RS_CHUNKS>), nBlocks, RS_THREADS, stream, kin, n, shift, s.hist, nBlocks);
        scan_exclusive_u32(s.hist, s.hist, 256u * nBlocks, nullptr, s.scan, stream);
        if (small) PR_LAUNCH_GRID((k_rs_scatter<K, RS_CHUNKS_SMALL>), nBlocks, RS_THREADS, stream, kin, iota ? nullptr : vin, kout, vout, n, shift,
                                  s.hist, nBlocks);
        else PR_LAUNCH_GRID((k_rs_scatter<K, RS_CHUNKS>), nBlocks, RS_THREADS, stream, kin, iota ? nullptr : vin, kout, vout, n, shift,
                       s.hist, nBlocks);
        iota = false;
        inA = !inA;
    }
    return !inA;
}

// ---------------------------------------------------------------- small sort (one block)
// Sorts up to SMALL_SORT_MAX unique 64-bit keys (full-key order) with their element indices as values,
// in shared memory with a bitonic network: the latency path for the few candidate pairs a TOI round or
// a quiet regular step produces (the tiled radix sort above needs 3 launches per 8-bit pass).
constexpr unsigned SMALL_SORT_MAX = 2048;
__global__ void k_sort_pairs_small(const unsigned long long* __restrict__ keysIn, unsigned long long* __restrict__ keysOut,
                                   unsigned* __restrict__ valsOut, unsigned n)
{
    PR_PDL_ENTER();
    __shared__ unsigned long long k[SMALL_SORT_MAX];
    __shared__ unsigned v[SMALL_SORT_MAX];
    for (unsigned i = threadIdx.x; i < SMALL_SORT_MAX; i += blockDim.x) {
        k[i] = (i < n) ? keysIn[i] : ~0ull;
        v[i] = i;
    }
    __syncthreads();
    for (unsigned size = 2; size <= SMALL_SORT_MAX; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = threadIdx.x; t < SMALL_SORT_MAX / 2; t += blockDim.x) {
                const unsigned lo = 2u * t - (t & (stride - 1u));
                const unsigned hi = lo + stride;
                const bool asc = ((lo & size) == 0u);
                const unsigned long long a = k[lo], b = k[hi];
                if ((a > b) == asc) { k[lo] = b; k[hi] = a; const unsigned tv = v[lo]; v[lo] = v[hi]; v[hi] = tv; }
            }
            __syncthreads();
        }
    }
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) { keysOut[i] = k[i]; valsOut[i] = v[i]; }
}
inline void sort_pairs_small(const unsigned long long* keysIn, unsigned long long* keysOut, unsigned* valsOut, unsigned n,
                             cudaStream_t stream)
{
    if (n) PR_LAUNCH_GRID(k_sort_pairs_small, 1, 1024, stream, keysIn, keysOut, valsOut, n);
}

#else // ---------------------------------------------------------------- host emulation
} // namespace pr
#include <vector>
namespace pr {

inline void scan_exclusive_u32(const unsigned* in, unsigned* out, unsigned n, unsigned* totalOut,
                               unsigned*, pr_stream_t)
{
    unsigned acc = 0;
    for (unsigned i = 0; i < n; ++i) { const unsigned v = in[i]; out[i] = acc; acc += v; }
    if (totalOut) *totalOut = acc;
    count_launch();
}
inline size_t scan_scratch_elems(unsigned n) { return n / 2048 + 2; }
struct RadixScratch { unsigned* hist = nullptr; unsigned* scan = nullptr; size_t histCap = 0, scanCap = 0; };
inline unsigned rs_blocks(unsigned n) { return (n + 4095) / 4096; }
template <class K>
inline bool radix_sort_pairs(K* keysA, unsigned* valsA, K* keysB, unsigned* valsB, unsigned n, int beginBit,
                             int endBit, bool valsAreIota, RadixScratch&, pr_stream_t)
{
    if (n == 0) return false;
    (void)keysB; (void)valsB;
    std::vector<unsigned> idx(n);
    for (unsigned i = 0; i < n; ++i) idx[i] = i;
    const int nbits = ((endBit - beginBit + 7) / 8) * 8;
    const K mask = (nbits >= static_cast<int>(sizeof(K) * 8)) ? ~K(0) : ((K(1) << nbits) - 1);
    std::stable_sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) {
        return ((keysA[a] >> beginBit) & mask) < ((keysA[b] >> beginBit) & mask);
    });
    std::vector<K> k(n); std::vector<unsigned> v(n);
    for (unsigned i = 0; i < n; ++i) { k[i] = keysA[idx[i]]; v[i] = valsAreIota ? idx[i] : valsA[idx[i]]; }
    for (unsigned i = 0; i < n; ++i) { keysA[i] = k[i]; valsA[i] = v[i]; }
    count_launch();
    return false;
}

constexpr unsigned SMALL_SORT_MAX = 2048;
inline void sort_pairs_small(const unsigned long long* keysIn, unsigned long long* keysOut, unsigned* valsOut, unsigned n,
                             pr_stream_t)
{
    std::vector<unsigned> idx(n);
    for (unsigned i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return keysIn[a] < keysIn[b]; });
    std::vector<unsigned long long> k(n);
    for (unsigned i = 0; i < n; ++i) k[i] = keysIn[idx[i]];
    for (unsigned i = 0; i < n; ++i) { keysOut[i] = k[i]; valsOut[i] = idx[i]; }
    if (n) count_launch();
}
#endif

} // namespace pr
#endif
