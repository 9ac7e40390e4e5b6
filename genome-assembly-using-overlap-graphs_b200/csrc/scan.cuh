// Exclusive prefix sum over device arrays.
//   store(i, sum(f(0..i))) for i < n and store.total(n, sum(f(0..n))); f is a device functor (a plain
//   array load, or a value computed on the fly so that no count array is ever written), store is an
//   output functor (one array, or the two components of a pair scan split over two arrays).
//   n <= kScanSmall : one CTA, one launch.
//   larger          : tile sums -> scan of tile sums (one CTA) -> apply   (2 reads + 1 write).
// Global accesses are coalesced: a CTA evaluates f and stores results in STRIPED order (consecutive
// threads touch consecutive elements) and moves the tile through shared memory to the BLOCKED order
// (16 consecutive elements per thread) the sequential part of the scan needs.
#pragma once
#include "common.cuh"

namespace ovl {

constexpr int kScanThreads = 256;
constexpr int64_t kScanSmallTiles = 4;                // up to 4 tiles: one launch beats three

struct I64x2 { int64_t x, y; };
__host__ __device__ __forceinline__ I64x2 operator+(I64x2 a, I64x2 b) { return I64x2{a.x + b.x, a.y + b.y}; }
__host__ __device__ __forceinline__ I64x2 operator-(I64x2 a, I64x2 b) { return I64x2{a.x - b.x, a.y - b.y}; }

template <typename T> struct ScanTraits {
    static constexpr int items = 16;                  // per thread
    __device__ static __forceinline__ T zero() { return (T)0; }
    __device__ static __forceinline__ T shfl_up(T v, int d) { return __shfl_up_sync(kFull, v, d); }
};
template <> struct ScanTraits<I64x2> {
    static constexpr int items = 8;
    __device__ static __forceinline__ I64x2 zero() { return I64x2{0, 0}; }
    __device__ static __forceinline__ I64x2 shfl_up(I64x2 v, int d) {
        return I64x2{__shfl_up_sync(kFull, v.x, d), __shfl_up_sync(kFull, v.y, d)};
    }
};
template <typename T> __host__ __device__ constexpr int scan_tile() { return kScanThreads * ScanTraits<T>::items; }
// shared-memory slot of tile element j: one pad element per 16, so that the blocked accesses
// (stride 16 elements across a warp) spread over the banks
__device__ __forceinline__ int scan_slot(int j) { return j + (j >> 4); }
template <typename T> __host__ __device__ constexpr int scan_smem_elems() { return scan_tile<T>() + scan_tile<T>() / 16 + 1; }

template <typename TI>
struct LoadArray {
    const TI* p;
    __device__ __forceinline__ TI operator()(int64_t i) const { return p[i]; }
};
template <typename TO>
struct StoreArray {
    TO* p;
    __device__ __forceinline__ void operator()(int64_t i, TO v) const { p[i] = v; }
};
// a pair scan whose two running sums go to two arrays
struct StoreSplit {
    int64_t* a;
    int64_t* b;
    __device__ __forceinline__ void operator()(int64_t i, I64x2 v) const { a[i] = v.x; b[i] = v.y; }
};

// copies[a[p]] * copies[b[p]]: the number of edges pair p expands to (overlapGraphs.py:55-57)
struct CopyProduct {
    const int32_t* pair_a;
    const int32_t* pair_b;
    const int32_t* copies;
    __device__ __forceinline__ int64_t operator()(int64_t p) const {
        return (int64_t)copies[pair_a[p]] * (int64_t)copies[pair_b[p]];
    }
};

// 1 when edge row e is kept by the score filter of the all-pairs builders (overlapGraphs.py:225, :347)
struct EdgeKept {
    const int4* edges;
    int32_t min_weight;
    __device__ __forceinline__ int64_t operator()(int64_t e) const { return edges[e].z >= min_weight ? 1 : 0; }
};

// exclusive scan of one value per thread across the CTA; *total = CTA sum.  smem: [warps + 1]
template <typename TO>
__device__ __forceinline__ TO block_exclusive_scan(TO v, TO* total, TO* smem) {
    using TR = ScanTraits<TO>;
    TO inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        TO o = TR::shfl_up(inc, d);
        if ((int)lane_id() >= d) inc = inc + o;
    }
    int w = threadIdx.x >> 5;
    if (lane_id() == 31) smem[w] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        TO run = TR::zero();
        for (int i = 0; i < kScanThreads / 32; ++i) { TO t = smem[i]; smem[i] = run; run = run + t; }
        smem[kScanThreads / 32] = run;
    }
    __syncthreads();
    TO excl = (inc - v) + smem[w];
    *total = smem[kScanThreads / 32];
    __syncthreads();                                   // smem is reused by the caller's next round
    return excl;
}

template <typename F, typename TO>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(F f, int64_t n, TO* __restrict__ tile_sums) {
    using TR = ScanTraits<TO>;
    __shared__ TO sm[kScanThreads / 32 + 1];
    const int64_t base = (int64_t)blockIdx.x * scan_tile<TO>();
    TO s = TR::zero();
#pragma unroll
    for (int i = 0; i < TR::items; ++i) {
        int64_t idx = base + i * kScanThreads + threadIdx.x;           // striped: coalesced
        if (idx < n) s = s + (TO)f(idx);
    }
    TO total;
    block_exclusive_scan<TO>(s, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// One tile: striped evaluation of f -> blocked running sums -> striped stores.  Returns the tile total.
template <typename F, typename S, typename TO>
__device__ __forceinline__ TO scan_one_tile(const F& f, const S& store, int64_t base, int64_t n, TO carry,
                                            TO* tile /*[scan_smem_elems]*/, TO* sm /*[warps + 1]*/) {
    using TR = ScanTraits<TO>;
    constexpr int IT = TR::items;
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        int j = i * kScanThreads + threadIdx.x;
        int64_t idx = base + j;
        tile[scan_slot(j)] = idx < n ? (TO)f(idx) : TR::zero();
    }
    __syncthreads();                                   // every input of the tile is read before any output is written
    TO v[IT];
    TO s = TR::zero();
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        v[i] = tile[scan_slot(threadIdx.x * IT + i)];
        s = s + v[i];
    }
    TO total;
    TO excl = block_exclusive_scan<TO>(s, &total, sm) + carry;
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        tile[scan_slot(threadIdx.x * IT + i)] = excl;
        excl = excl + v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        int j = i * kScanThreads + threadIdx.x;
        int64_t idx = base + j;
        if (idx < n) store(idx, tile[scan_slot(j)]);
    }
    __syncthreads();                                   // the tile buffer is reused
    return total;
}

// one CTA: exclusive scan of f(0..n), looping over tiles
template <typename F, typename S, typename TO>
__global__ void __launch_bounds__(kScanThreads) scan_single_cta(F f, S store, int64_t n, bool write_total) {
    __shared__ TO sm[kScanThreads / 32 + 1];
    __shared__ TO tile[scan_smem_elems<TO>()];
    TO carry = ScanTraits<TO>::zero();
    for (int64_t start = 0; start < n; start += scan_tile<TO>())
        carry = carry + scan_one_tile<F, S, TO>(f, store, start, n, carry, tile, sm);
    if (write_total && threadIdx.x == 0) store(n, carry);
}

template <typename F, typename S, typename TO>
__global__ void __launch_bounds__(kScanThreads) scan_apply(F f, S store, int64_t n, const TO* __restrict__ tile_offsets) {
    __shared__ TO sm[kScanThreads / 32 + 1];
    __shared__ TO tile[scan_smem_elems<TO>()];
    const int64_t base = (int64_t)blockIdx.x * scan_tile<TO>();
    TO off = tile_offsets[blockIdx.x];
    TO total = scan_one_tile<F, S, TO>(f, store, base, n, off, tile, sm);
    if (base + scan_tile<TO>() >= n && threadIdx.x == 0) store(n, off + total);     // grand total, by the last tile
}

inline size_t scan_workspace_bytes(int64_t n, size_t elem) {
    int64_t tile = kScanThreads * (elem > 8 ? 8 : 16);
    int64_t nb = (n + tile - 1) / tile;
    return (size_t)(nb > 0 ? nb : 1) * elem;
}

// store(i, .) for i in [0, n] (n + 1 results).  When f reads an array that the store aliases (in-place
// scan) every CTA reads all inputs of a tile before it writes any output of that tile, and slot n lies
// outside f's range.  Returns the number of kernels launched through *launches.
template <typename F, typename S, typename TO>
inline cudaError_t exclusive_scan_to(F f, S store, int64_t n, void* ws, cudaStream_t st, int* launches = nullptr) {
    if (n < 0) n = 0;
    const int64_t tile = scan_tile<TO>();
    if (n <= kScanSmallTiles * tile) {
        scan_single_cta<F, S, TO><<<1, kScanThreads, 0, st>>>(f, store, n, true);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    int64_t nb = (n + tile - 1) / tile;
    TO* sums = reinterpret_cast<TO*>(ws);
    scan_tile_sums<F, TO><<<(unsigned)nb, kScanThreads, 0, st>>>(f, n, sums);
    scan_single_cta<LoadArray<TO>, StoreArray<TO>, TO><<<1, kScanThreads, 0, st>>>(LoadArray<TO>{sums}, StoreArray<TO>{sums}, nb, false);
    scan_apply<F, S, TO><<<(unsigned)nb, kScanThreads, 0, st>>>(f, store, n, sums);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

// out: n + 1 elements
template <typename F, typename TO>
inline cudaError_t exclusive_scan(F f, TO* out, int64_t n, void* ws, cudaStream_t st, int* launches = nullptr) {
    return exclusive_scan_to<F, StoreArray<TO>, TO>(f, StoreArray<TO>{out}, n, ws, st, launches);
}

}  // namespace ovl
