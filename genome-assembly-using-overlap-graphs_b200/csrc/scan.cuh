// Exclusive prefix sum over device arrays.
//   out has n+1 entries: out[i] = sum(f(0..i)), out[n] = total;  f(i) is a device functor (a
//   plain array load, or a value computed on the fly so that no count array is ever written).
//   n <= kScanSmall : one CTA, one launch.
//   larger          : reduce -> scan of tile sums (one CTA) -> apply   (2 reads + 1 write).
#pragma once
#include "common.cuh"

namespace ovl {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096 elements per CTA
constexpr int64_t kScanSmall = 4 * kScanTile;         // up to 16,384 elements: one launch beats three

template <typename TI>
struct LoadArray {
    const TI* p;
    __device__ __forceinline__ TI operator()(int64_t i) const { return p[i]; }
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

template <typename TO>
__device__ __forceinline__ TO block_exclusive_scan(TO v, TO* total, TO* smem /*[8+1]*/) {
    // inclusive warp scan of per-thread sums, then across the 8 warps
    TO inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        TO o = __shfl_up_sync(kFull, inc, d);
        if ((int)lane_id() >= d) inc += o;
    }
    int w = threadIdx.x >> 5;
    if (lane_id() == 31) smem[w] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        TO run = 0;
        for (int i = 0; i < kScanThreads / 32; ++i) { TO t = smem[i]; smem[i] = run; run += t; }
        smem[kScanThreads / 32] = run;
    }
    __syncthreads();
    TO excl = inc - v + smem[w];
    *total = smem[kScanThreads / 32];
    __syncthreads();                                   // smem is reused by the caller's next round
    return excl;
}

template <typename F, typename TO>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(F f, int64_t n, TO* __restrict__ tile_sums) {
    __shared__ TO sm[kScanThreads / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    TO s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t idx = base + i;
        if (idx < n) s += (TO)f(idx);
    }
    TO total;
    block_exclusive_scan<TO>(s, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one CTA: exclusive scan of f(0..n) into out[0..n] (out[n] = total), looping over tiles
template <typename F, typename TO>
__global__ void __launch_bounds__(kScanThreads) scan_single_cta(F f, int64_t n, TO* out, bool write_total) {
    __shared__ TO sm[kScanThreads / 32 + 1];
    TO carry = 0;
    for (int64_t start = 0; start < n; start += kScanTile) {
        int64_t base = start + (int64_t)threadIdx.x * kScanItems;
        TO v[kScanItems];
        TO s = 0;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            int64_t idx = base + i;
            v[i] = idx < n ? (TO)f(idx) : (TO)0;
            s += v[i];
        }
        TO total;
        TO excl = block_exclusive_scan<TO>(s, &total, sm) + carry;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            int64_t idx = base + i;
            if (idx < n) out[idx] = excl;
            excl += v[i];
        }
        carry += total;
    }
    if (write_total && threadIdx.x == 0) out[n] = carry;
}

template <typename F, typename TO>
__global__ void __launch_bounds__(kScanThreads) scan_apply(F f, int64_t n, const TO* __restrict__ tile_offsets, TO* out) {
    __shared__ TO sm[kScanThreads / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    TO v[kScanItems];
    TO s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t idx = base + i;
        v[i] = idx < n ? (TO)f(idx) : (TO)0;
        s += v[i];
    }
    TO total;
    TO excl = block_exclusive_scan<TO>(s, &total, sm) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t idx = base + i;
        if (idx < n) out[idx] = excl;
        excl += v[i];
        if (idx == n - 1) out[n] = excl;     // grand total
    }
}

inline size_t scan_workspace_bytes(int64_t n, size_t elem) {
    int64_t nb = (n + kScanTile - 1) / kScanTile;
    return (size_t)(nb > 0 ? nb : 1) * elem;
}

// out: n+1 elements.  When f reads an array that aliases out (in-place scan) every thread
// reads all of its elements before it writes any, and out[n] lies outside f's range.
// Returns the number of kernels launched through *launches (bookkeeping for the bench).
template <typename F, typename TO>
inline cudaError_t exclusive_scan(F f, TO* out, int64_t n, void* ws, cudaStream_t st, int* launches = nullptr) {
    if (n <= 0) {
        return cudaMemsetAsync(out, 0, sizeof(TO), st);
    }
    if (n <= kScanSmall) {
        scan_single_cta<F, TO><<<1, kScanThreads, 0, st>>>(f, n, out, true);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    int64_t nb = (n + kScanTile - 1) / kScanTile;
    TO* sums = reinterpret_cast<TO*>(ws);
    scan_tile_sums<F, TO><<<(unsigned)nb, kScanThreads, 0, st>>>(f, n, sums);
    scan_single_cta<LoadArray<TO>, TO><<<1, kScanThreads, 0, st>>>(LoadArray<TO>{sums}, nb, sums, false);
    scan_apply<F, TO><<<(unsigned)nb, kScanThreads, 0, st>>>(f, n, sums, out);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

}  // namespace ovl
