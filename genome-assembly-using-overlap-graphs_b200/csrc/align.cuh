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

// K8, row-per-thread variant for queries up to 1,024 symbols (reads and typical contigs): thread t owns
// query row i = t + 1 for the whole sweep.  On diagonal d it computes cell (i, j = d - i); its "left"
// H[i][j-1] is its own previous output, its "up" H[i-1][j] is the left neighbour's previous output (one
// __shfl_up_sync) and its "diag" H[i-1][j-1] the one before that.  Only warp boundaries go through
// shared memory, double buffered, so a diagonal costs one block barrier (none for a single warp).
// tb is diagonal-major here: byte (d, i) at tb[d*(n+1) + i] -- consecutive threads, consecutive bytes.
__device__ __forceinline__ void local_align_rows_body(const int32_t* __restrict__ q, int n,
                                                      const int32_t* __restrict__ ref, int m,
                                                      int64_t match, int64_t mismatch, int64_t indel,
                                                      int8_t* __restrict__ tb,
                                                      int32_t* __restrict__ result, uint8_t* __restrict__ ops) {
    __shared__ int32_t edge[2][kAlignThreads / 32];
    __shared__ int32_t s_best[kAlignThreads / 32], s_bi[kAlignThreads / 32], s_bj[kAlignThreads / 32];
    const int t = threadIdx.x, i = t + 1, wid = t >> 5;
    const unsigned lane = lane_id();
    const bool row_ok = i <= n;
    const int S = n + 1;
    const int32_t qi = row_ok ? q[i - 1] : -1;
    int32_t out_prev = 0;        // my cell on the previous diagonal  (H[i][j-1], 0 outside the matrix)
    int32_t up_prev = 0;         // the neighbour's cell two diagonals ago (H[i-1][j-1])
    int32_t best = 0, bj = 0;
    if (t < kAlignThreads / 32) { edge[0][t] = 0; edge[1][t] = 0; s_best[t] = 0; s_bi[t] = 0; s_bj[t] = 0; }
    __syncthreads();
    for (int d = 2; d <= n + m; ++d) {
        int32_t up = __shfl_up_sync(kFull, out_prev, 1);
        if (lane == 0) up = wid == 0 ? 0 : edge[(d - 1) & 1][wid - 1];     // row i-1 lives in the previous warp (or is row 0)
        const int j = d - i;
        int32_t v = 0;
        if (row_ok && j >= 1 && j <= m) {
            int64_t dg = (int64_t)up_prev + (qi == ref[j - 1] ? match : mismatch);
            int64_t u = (int64_t)up + indel;
            int64_t lf = (int64_t)out_prev + indel;
            int8_t dir;
            if (dg >= u && dg >= lf && dg >= 0) { v = (int32_t)dg; dir = 1; }
            else if (u >= lf && u >= 0)         { v = (int32_t)u;  dir = 2; }
            else if (lf >= 0)                   { v = (int32_t)lf; dir = 3; }
            else                                { v = 0;           dir = 0; }
            tb[(size_t)d * S + i] = (int8_t)(dir | (v > 0 ? 4 : 0));
            if (v > best) { best = v; bj = j; }            // j ascends along my row: first strict maximum
        }
        up_prev = up;
        out_prev = v;
        if (lane == 31) edge[d & 1][wid] = v;
        if (blockDim.x > 32) __syncthreads();
    }
    // block arg-max: value, then smaller i, then smaller j (row-major first maximum, aligners.py:135-137)
    int32_t bi = i;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        int32_t ob = __shfl_xor_sync(kFull, best, off), oi = __shfl_xor_sync(kFull, bi, off), oj = __shfl_xor_sync(kFull, bj, off);
        if (ob > best || (ob == best && ob > 0 && (oi < bi || (oi == bi && oj < bj)))) { best = ob; bi = oi; bj = oj; }
    }
    if (lane == 0) { s_best[wid] = best; s_bi[wid] = bi; s_bj[wid] = bj; }
    __syncthreads();
    if (t == 0) {
        for (int w = 1; w < kAlignThreads / 32; ++w) {
            int32_t ob = s_best[w], oi = s_bi[w], oj = s_bj[w];
            if (ob > best || (ob == best && ob > 0 && (oi < bi || (oi == bi && oj < bj)))) { best = ob; bi = oi; bj = oj; }
        }
        if (best == 0) { bi = 0; bj = 0; }                 // nothing beat the initial (0, 0), aligners.py:113-114
        int ii = bi, jj = bj, L = 0;
        while (ii > 0 && jj > 0) {
            int8_t x = tb[(size_t)(ii + jj) * S + ii];
            if (!(x & 4)) break;
            int dir = x & 3;
            if (dir == 0) break;
            ops[L++] = (uint8_t)dir;
            if (dir == 1) { --ii; --jj; } else if (dir == 2) { --ii; } else { --jj; }
        }
        result[0] = best;
        result[1] = jj;
        result[2] = bj;
        result[3] = L;
        result[4] = bi;
    }
}

__global__ void __launch_bounds__(kAlignThreads) local_align_rows_kernel(const int32_t* __restrict__ q, int n,
                                                                         const int32_t* __restrict__ ref, int m,
                                                                         int64_t match, int64_t mismatch, int64_t indel,
                                                                         int8_t* __restrict__ tb,
                                                                         int32_t* __restrict__ result, uint8_t* __restrict__ ops) {
    local_align_rows_body(q, n, ref, m, match, mismatch, indel, tb, result, ops);
}

// K8 batch: one CTA per query, all queries against windows of ONE reference that is read from L2 by every
// CTA (performanceMeasures.py:219-221 aligns every contig of an assembly to the same genome, one call
// each: 148+ of them now run side by side).  Query x: symbols queries[q_off[x] .. q_off[x+1]), reference
// window [ref_start[x], ref_start[x] + ref_len[x]), traceback bytes at tb + tb_off[x], op list at
// ops + ops_off[x], result row results + 8 x.  Every query has at most blockDim.x symbols.
__global__ void __launch_bounds__(kAlignThreads) local_align_batch_kernel(const int32_t* __restrict__ queries,
                                                                          const int64_t* __restrict__ q_off,
                                                                          const int32_t* __restrict__ reference,
                                                                          const int32_t* __restrict__ ref_start,
                                                                          const int32_t* __restrict__ ref_len,
                                                                          int64_t match, int64_t mismatch, int64_t indel,
                                                                          int8_t* __restrict__ tb, const int64_t* __restrict__ tb_off,
                                                                          int32_t* __restrict__ results,
                                                                          uint8_t* __restrict__ ops, const int64_t* __restrict__ ops_off) {
    const int x = blockIdx.x;
    const int64_t q0 = q_off[x];
    const int n = (int)(q_off[x + 1] - q0);
    const int m = ref_len[x];
    int32_t* result = results + 8 * (size_t)x;
    if (n == 0 || m == 0) {                                  // empty query or window: score 0 at (0, 0), aligners.py:113-114
        if (threadIdx.x < 8) result[threadIdx.x] = 0;
        return;
    }
    local_align_rows_body(queries + q0, n, reference + ref_start[x], m, match, mismatch, indel, tb + tb_off[x], result,
                          ops + ops_off[x]);
}

}  // namespace ovl
