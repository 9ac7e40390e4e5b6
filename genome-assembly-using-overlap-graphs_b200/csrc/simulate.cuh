// Seeded read simulator on the device: the reference's generators as a counter-based
// stream, so that every base of every read is an independent function of (seed, read, position) and the
// same reads come out of this kernel and of its NumPy mirror (synth.simulate_reads_counter) bit for bit.
//   generateErrorFreeReads.py:38-50  uniform start on a LINEAR genome, reads truncated at the genome end
//   generateErrorProneReads.py:17-28 every base replaced with probability p by one of the 3 other bases
// (The reference itself is un-seeded -- Python `random` and Numba's RNG -- so no stream of its own exists
// to reproduce; what is kept is the distribution.)
#pragma once
#include "common.cuh"

namespace ovl {

// splitmix64 finaliser over (seed, stream, index)
__host__ __device__ __forceinline__ uint64_t sim_hash(uint64_t seed, uint64_t stream, uint64_t index) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (stream + 1) + 0xD1B54A32D192ED03ull * index;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// read i: start = floor(hi32(h) * G / 2^32), length = min(read_len, G - start)
__global__ void __launch_bounds__(256) sim_starts_kernel(int64_t n_reads, int64_t G, int32_t read_len, uint64_t seed,
                                                         int64_t* __restrict__ start, int64_t* __restrict__ len) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    uint64_t h = sim_hash(seed, 0, (uint64_t)i);
    int64_t s = (int64_t)(((h >> 32) * (uint64_t)G) >> 32);
    start[i] = s;
    len[i] = min((int64_t)read_len, G - s);
}

// one thread per 16 output bases of one read.  Base j of read i: the genome letter, replaced when
// hi32(h(seed, 1, i * read_len + j)) < error_thr by the letter (code + 1 + floor(lo32(h) * 3 / 2^32)) & 3
// in the order A, C, G, T.
__global__ void __launch_bounds__(256) sim_bases_kernel(const uint8_t* __restrict__ genome, int64_t n_reads, int32_t read_len,
                                                        uint32_t error_thr, uint64_t seed, const int64_t* __restrict__ start,
                                                        const int64_t* __restrict__ offsets, uint8_t* __restrict__ out) {
    const int chunks = (read_len + 15) / 16;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i = t / chunks;
    if (i >= n_reads) return;
    int c = (int)(t - i * chunks);
    int64_t o0 = offsets[i];
    int len = (int)(offsets[i + 1] - o0);
    int64_t s = start[i];
    for (int j = 16 * c; j < min(len, 16 * c + 16); ++j) {
        uint8_t b = genome[s + j];
        uint64_t h = sim_hash(seed, 1, (uint64_t)i * (uint64_t)read_len + (uint64_t)j);
        if ((uint32_t)(h >> 32) < error_thr) {
            uint32_t code = b == 'A' ? 0u : b == 'C' ? 1u : b == 'G' ? 2u : 3u;
            uint32_t shift = 1u + (uint32_t)(((h & 0xffffffffull) * 3ull) >> 32);
            b = "ACGT"[(code + shift) & 3u];
        }
        out[o0 + j] = b;
    }
}

}  // namespace ovl
