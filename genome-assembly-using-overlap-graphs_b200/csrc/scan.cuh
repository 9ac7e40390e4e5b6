// Exclusive prefix sum over device arrays (reduce -> scan of tile sums -> apply).
// out has n+1 entries: out[i] = sum(in[0..i)), out[n] = total.
#pragma once
#include "common.cuh"

namespace ovl {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096 elements per CTA

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
    return excl;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const TI* __restrict__ in, int64_t n,
                                                               TO* __restrict__ tile_sums) {
    __shared__ TO sm[kScanThreads / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    TO s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t idx = base + i;
        if (idx < n) s += (TO)in[idx];
    }
    TO total;
    block_exclusive_scan<TO>(s, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single CTA: in-place exclusive scan of the tile sums
template <typename TO>
__global__ void __launch_bounds__(kScanThreads) scan_sums_inplace(TO* __restrict__ sums, int64_t nb) {
    __shared__ TO sm[kScanThreads / 32 + 1];
    TO carry = 0;
    for (int64_t start = 0; start < nb; start += kScanTile) {
        int64_t base = start + (int64_t)threadIdx.x * kScanItems;
        TO v[kScanItems];
        TO s = 0;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            int64_t idx = base + i;
            v[i] = idx < nb ? sums[idx] : (TO)0;
            s += v[i];
        }
        TO total;
        TO excl = block_exclusive_scan<TO>(s, &total, sm) + carry;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            int64_t idx = base + i;
            if (idx < nb) sums[idx] = excl;
            excl += v[i];
        }
        carry += total;
        __syncthreads();
    }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(kScanThreads) scan_apply(const TI* in, int64_t n,
                                                           const TO* __restrict__ tile_offsets,
                                                           TO* out) {
    __shared__ TO sm[kScanThreads / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    TO v[kScanItems];
    TO s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t idx = base + i;
        v[i] = idx < n ? (TO)in[idx] : (TO)0;
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

// in: n elements, out: n+1 elements (may alias in only if sizeof(TI)==sizeof(TO) and the
// caller does not need out[n] to land outside in's allocation).
template <typename TI, typename TO>
inline cudaError_t exclusive_scan(const TI* in, TO* out, int64_t n, void* ws, cudaStream_t st) {
    if (n <= 0) {
        return cudaMemsetAsync(out, 0, sizeof(TO), st);
    }
    int64_t nb = (n + kScanTile - 1) / kScanTile;
    TO* sums = reinterpret_cast<TO*>(ws);
    scan_tile_sums<TI, TO><<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, sums);
    scan_sums_inplace<TO><<<1, kScanThreads, 0, st>>>(sums, nb);
    scan_apply<TI, TO><<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, sums, out);
    return cudaGetLastError();
}

}  // namespace ovl
