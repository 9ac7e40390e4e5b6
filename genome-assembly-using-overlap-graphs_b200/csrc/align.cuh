// K7: one pair with traceback -- the whole of aligners.py:27-76 for the single-pair drop-in
// (overlap_alignment returns the aligned strings as well as score/end).
//
// One CTA sweeps the anti-diagonals d = i + j; the cells of a diagonal are independent.
// Arithmetic follows the reference exactly: candidates in int64 (Numba types indel as int64),
// stored truncated to int32 (aligners.py:28, 35-48), tie order diag >= up >= left, first
// strict maximum over the last row (aligners.py:50-57), traceback walk while i > 0 and j > 0
// (aligners.py:63-76).  Sequences are int32 code points, so any alphabet works here.
#pragma once
#include "common.cuh"

namespace ovl {

constexpr int kAlignThreads = 1024;

// workspace layout (int32 units): diag[3][n+1] | last_row[m+1] ; then tb[(n+1)*(m+1)] bytes.
// SMEM = true keeps the three rolling anti-diagonals in shared memory (when 3*(n+1) ints fit).
template <bool SMEM>
__global__ void __launch_bounds__(kAlignThreads) align_pair_kernel(const int32_t* __restrict__ s, int n,
                                                                   const int32_t* __restrict__ t, int m,
                                                                   int64_t match, int64_t mismatch, int64_t indel,
                                                                   int32_t* __restrict__ diag_g, int32_t* __restrict__ last_row,
                                                                   int8_t* __restrict__ tb,
                                                                   int32_t* __restrict__ result, uint8_t* __restrict__ ops) {
    extern __shared__ int32_t diag_sh[];
    int32_t* diag = SMEM ? diag_sh : diag_g;
    const int W = m + 1;
    const int stride = n + 1;
    for (int j = threadIdx.x; j <= m; j += blockDim.x) last_row[j] = 0;     // n == 0: row 0 is all zero
    for (int i = threadIdx.x; i < 3 * stride; i += blockDim.x) diag[i] = 0;
    __syncthreads();
    for (int d = 2; d <= n + m; ++d) {
        int32_t* cur = diag + (d % 3) * stride;
        const int32_t* p1 = diag + ((d - 1) % 3) * stride;
        const int32_t* p2 = diag + ((d - 2) % 3) * stride;
        int ilo = max(1, d - m), ihi = min(n, d - 1);
        for (int i = ilo + (int)threadIdx.x; i <= ihi; i += blockDim.x) {
            int j = d - i;
            int64_t dg = (int64_t)p2[i - 1] + (s[i - 1] == t[j - 1] ? match : mismatch);
            int64_t up = (int64_t)p1[i - 1] + indel;
            int64_t lf = (int64_t)p1[i] + indel;
            int32_t v;
            int8_t dir;
            if (dg >= up && dg >= lf) { v = (int32_t)dg; dir = 0; }
            else if (up >= lf)        { v = (int32_t)up; dir = 1; }
            else                      { v = (int32_t)lf; dir = 2; }
            cur[i] = v;
            tb[(size_t)i * W + j] = dir;
            if (i == n) last_row[j] = v;
        }
        // boundary cells of this diagonal: dp[0][d] and dp[d][0] are zero
        if (threadIdx.x == 0) { cur[0] = 0; if (d <= n) cur[d] = 0; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int32_t best = last_row[0];          // j = 0 beats -inf (aligners.py:51-57)
        int bj = 0;
        for (int j = 1; j <= m; ++j) if (last_row[j] > best) { best = last_row[j]; bj = j; }
        int i = n, j = bj, L = 0;
        while (i > 0 && j > 0) {
            int8_t dir = tb[(size_t)i * W + j];
            ops[L++] = (uint8_t)dir;
            if (dir == 0) { --i; --j; } else if (dir == 1) { --i; } else { --j; }
        }
        result[0] = best;
        result[1] = bj;
        result[2] = L;
    }
}

// K8: Smith-Waterman local alignment with traceback -- the whole of aligners.py:106-162
// (local_alignment), the evaluation-side aligner that maps reads / contigs back to the genome.
// Same anti-diagonal sweep as K7.  Differences, all as in the reference: cells are floored at 0
// with the tie order diag >= up >= left >= "restart" (aligners.py:121-132), the best cell is the
// first strict maximum in row-major order (:135-137), the walk stops at a zero cell (:143-160).
// tb byte = direction (1 diag, 2 up, 3 left, 0 restart) | 4 if the cell value is > 0.
// The three rolling anti-diagonals live in shared memory when 3*(n+1) ints fit (SMEM = true: one
// shared-memory round trip per diagonal instead of an L2 one), else in the global workspace.
template <bool SMEM>
__global__ void __launch_bounds__(kAlignThreads) local_align_kernel(const int32_t* __restrict__ q, int n,
                                                                    const int32_t* __restrict__ ref, int m,
                                                                    int64_t match, int64_t mismatch, int64_t indel,
                                                                    int32_t* __restrict__ diag_g, int8_t* __restrict__ tb,
                                                                    int32_t* __restrict__ result, uint8_t* __restrict__ ops) {
    __shared__ int32_t s_best[kAlignThreads / 32];
    __shared__ int32_t s_bi[kAlignThreads / 32], s_bj[kAlignThreads / 32];
    extern __shared__ int32_t diag_s[];
    int32_t* diag = SMEM ? diag_s : diag_g;
    const int W = m + 1;
    const int stride = n + 1;
    for (int i = threadIdx.x; i < 3 * stride; i += blockDim.x) diag[i] = 0;
    int32_t best = 0, bi = 0, bj = 0;            // best_score starts at 0 at (0, 0), aligners.py:113-114
    __syncthreads();
    for (int d = 2; d <= n + m; ++d) {
        int32_t* cur = diag + (d % 3) * stride;
        const int32_t* p1 = diag + ((d - 1) % 3) * stride;
        const int32_t* p2 = diag + ((d - 2) % 3) * stride;
        int ilo = max(1, d - m), ihi = min(n, d - 1);
        for (int i = ilo + (int)threadIdx.x; i <= ihi; i += blockDim.x) {
            int j = d - i;
            int64_t dg = (int64_t)p2[i - 1] + (q[i - 1] == ref[j - 1] ? match : mismatch);
            int64_t up = (int64_t)p1[i - 1] + indel;
            int64_t lf = (int64_t)p1[i] + indel;
            int32_t v;
            int8_t dir;
            if (dg >= up && dg >= lf && dg >= 0) { v = (int32_t)dg; dir = 1; }
            else if (up >= lf && up >= 0)         { v = (int32_t)up; dir = 2; }
            else if (lf >= 0)                     { v = (int32_t)lf; dir = 3; }
            else                                  { v = 0;           dir = 0; }
            cur[i] = v;
            tb[(size_t)i * W + j] = (int8_t)(dir | (v > 0 ? 4 : 0));
            // first strict maximum in row-major order == max value, then smallest i, then smallest j
            if (v > best || (v == best && v > 0 && (i < bi || (i == bi && j < bj)))) { best = v; bi = i; bj = j; }
        }
        if (threadIdx.x == 0) { cur[0] = 0; if (d <= n) cur[d] = 0; }
        __syncthreads();
    }
    // block arg-max with the same tie rule (warps beyond blockDim hold the neutral (0, 0, 0))
    if (threadIdx.x < kAlignThreads / 32) { s_best[threadIdx.x] = 0; s_bi[threadIdx.x] = 0; s_bj[threadIdx.x] = 0; }
    __syncthreads();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        int32_t ob = __shfl_xor_sync(kFull, best, off), oi = __shfl_xor_sync(kFull, bi, off), oj = __shfl_xor_sync(kFull, bj, off);
        if (ob > best || (ob == best && ob > 0 && (oi < bi || (oi == bi && oj < bj)))) { best = ob; bi = oi; bj = oj; }
    }
    if (lane_id() == 0) { s_best[threadIdx.x >> 5] = best; s_bi[threadIdx.x >> 5] = bi; s_bj[threadIdx.x >> 5] = bj; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kAlignThreads / 32; ++w) {
            int32_t ob = s_best[w], oi = s_bi[w], oj = s_bj[w];
            if (ob > best || (ob == best && ob > 0 && (oi < bi || (oi == bi && oj < bj)))) { best = ob; bi = oi; bj = oj; }
        }
        int i = bi, j = bj, L = 0;
        while (i > 0 && j > 0) {
            int8_t t = tb[(size_t)i * W + j];
            if (!(t & 4)) break;                 // dp[i][j] > 0 fails (aligners.py:143)
            int dir = t & 3;
            if (dir == 0) break;                 // aligners.py:159-160
            ops[L++] = (uint8_t)dir;
            if (dir == 1) { --i; --j; } else if (dir == 2) { --i; } else { --j; }
        }
        result[0] = best;
        result[1] = j;                           // start_pos (aligners.py:163)
        result[2] = bj;                          // end_pos   (aligners.py:164)
        result[3] = L;
        result[4] = bi;
    }
}

}  // namespace ovl
