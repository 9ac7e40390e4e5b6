// Order-sensitive fingerprint of an edge list (bench / multi-GPU self-checks).
//   H = sum over rows of mix(global row index, row)   (mod 2^64)
// The global row index is folded into every term, so a row in the wrong place, a swapped pair of
// rows or a row written twice changes H, while shards can still be hashed independently and
// added.  The host mirror of this function lives in engine.edge_hash_numpy().
#pragma once
#include "common.cuh"

namespace ovl {

__host__ __device__ __forceinline__ uint64_t edge_row_mix(uint64_t idx, uint32_t a, uint32_t b, uint32_t w, uint32_t e) {
    uint64_t z = idx * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    z ^= (uint64_t)a | ((uint64_t)b << 32);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z ^= (uint64_t)w | ((uint64_t)e << 32);
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) edge_hash_kernel(const int4* __restrict__ edges, int64_t E, int64_t first_row,
                                                        unsigned long long* __restrict__ accum) {
    uint64_t h = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
        int4 r = edges[i];
        h += edge_row_mix((uint64_t)(first_row + i), (uint32_t)r.x, (uint32_t)r.y, (uint32_t)r.z, (uint32_t)r.w);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) h += __shfl_xor_sync(kFull, h, d);
    if (lane_id() == 0 && h != 0) atomicAdd(accum, (unsigned long long)h);
}

}  // namespace ovl
