// K4/K5: the suffix-prefix overlap DP (replaces aligners.py:27-57) as an integer wavefront.
//
//   H[0][j] = H[i][0] = 0
//   H[i][j] = max(H[i-1][j-1] + (s[i-1]==t[j-1] ? match : mismatch), H[i-1][j] + indel, H[i][j-1] + indel)
//   score = max_j H[n][j] (first maximum, j = 0 included), end = that j
//
// Work decomposition.  A group of G lanes (G | 32) owns one *couple* of candidate pairs; lane
// r holds T consecutive DP columns [rT, rT+T) of the current row in registers and the group
// sweeps the rows as a wavefront: at step k lane r computes row k-r, taking its left
// neighbour's last column from the previous step with one __shfl_up_sync.  G*T >= longest t.
//
// Packed arithmetic.  The two pairs of a couple live in the two 16-bit halves of every
// register and are advanced together by the DPX instructions VIADDMNMX.S16x2
// (__viaddmax_s16x2): per 32-bit lane op two DP cells.  To keep the per-cell instruction
// count at four the recurrence is re-based:
//     G[i][j] = H[i][j] - i*base + beta,   base = min(match, mismatch)
//   diag:  G[i-1][j-1] + sc,  sc = score - base  in {0, |match - mismatch|}   (>= 0)
//   up:    G[i-1][j]   + (indel - base)
//   left:  G[i][j-1]   + indel
// sc for both halves comes from ONE byte permute: per row the two query bases select two
// 4-entry byte tables (lutA, lutB); per column the selector register holds the two target
// bases (PRMT picks lutA[tA] into the low half, lutB[tB] into the high half, zero bytes via
// the sign-replicate selector on a non-negative byte).  Because every G is in [0, 32767] the
// diagonal add is a plain 32-bit add with no carry between halves; it is issued as IMAD so
// it runs on the FMA pipe while PRMT and the two VIADDMNMX run on the ALU pipe.
// indel = -2^31 (the reference default: gapless) and any indel too negative to ever win are
// clamped to -(range+1), which is exact (see ovl_overlap_dp in ovl.cu for the bounds).
//
// A 32-bit variant (one pair per group, VIADDMNMX on s32, compare+select for sc) covers
// scoring schemes or read lengths whose range does not fit 16 bits.
#pragma once
#include "common.cuh"

namespace ovl {

struct DpParams {
    int32_t eqv;        // match - base      (>= 0)
    int32_t nev;        // mismatch - base   (>= 0), one of eqv/nev is 0
    int32_t base;       // min(match, mismatch)
    int32_t beta;       // bias so that every G >= 0
    int32_t gu;         // effective (indel - base), clamped
    int32_t gl;         // effective indel, clamped
    uint32_t one;       // == 1, a runtime value so the diagonal add stays an IMAD
};

constexpr int kDpThreads = 128;
#ifndef OVL_DP_MINB
#define OVL_DP_MINB 4          // resident CTAs per SM the register allocator must allow
#endif

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t imad_add(uint32_t a, uint32_t one, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t base_code(const uint32_t* row, int i) {
    return (row[i >> 4] >> ((i & 15) * 2)) & 3u;
}
__device__ __forceinline__ uint32_t pack2(int v) { return ((uint32_t)v & 0xffffu) * 0x10001u; }
__device__ __forceinline__ int half_lo(uint32_t x) { return (int)(int16_t)(x & 0xffffu); }
__device__ __forceinline__ int half_hi(uint32_t x) { return (int)x >> 16; }

// PK = true : two pairs per group (16-bit halves).  PK = false: one pair per group (32 bit).
template <int G, int T, bool PK>
__global__ void __launch_bounds__(kDpThreads, OVL_DP_MINB) overlap_dp_kernel(const uint32_t* __restrict__ packed, int row_words,
                                                                const int32_t* __restrict__ len,
                                                                const int32_t* __restrict__ pair_a,
                                                                const int32_t* __restrict__ pair_b, int64_t P,
                                                                DpParams prm,
                                                                int32_t* __restrict__ score_out,
                                                                int32_t* __restrict__ end_out) {
    constexpr int PAIRS = PK ? 2 : 1;
    constexpr int GROUPS_PER_WARP = 32 / G;
    constexpr int GROUPS_PER_CTA = (kDpThreads / 32) * GROUPS_PER_WARP;
    extern __shared__ uint32_t smem[];                 // [GROUPS_PER_CTA][2*PAIRS][row_words]

    const unsigned lane = lane_id();
    const int r = lane % G;                            // lane within the group
    const int gib = (threadIdx.x >> 5) * GROUPS_PER_WARP + lane / G;   // group within the CTA
    const int64_t grp = (int64_t)blockIdx.x * GROUPS_PER_CTA + gib;
    const int64_t p0 = grp * PAIRS;

    // ---- stage the packed reads of this group's pairs in shared memory (128-bit copies)
    int32_t n[PAIRS], m[PAIRS];
    uint32_t* rows = smem + (size_t)gib * (2 * PAIRS) * row_words;
    const int rw4 = row_words >> 2;
#pragma unroll
    for (int h = 0; h < PAIRS; ++h) {
        int64_t p = p0 + h;
        bool live = p < P;
        int32_t a = live ? pair_a[p] : 0, b = live ? pair_b[p] : 0;
        n[h] = live ? len[a] : 0;
        m[h] = live ? len[b] : 0;
        const uint4* sa = reinterpret_cast<const uint4*>(packed + (size_t)a * row_words);
        const uint4* sb = reinterpret_cast<const uint4*>(packed + (size_t)b * row_words);
        uint4* ds = reinterpret_cast<uint4*>(rows + (size_t)(2 * h) * row_words);
        uint4* dt = reinterpret_cast<uint4*>(rows + (size_t)(2 * h + 1) * row_words);
        for (int i = r; i < rw4; i += G) { ds[i] = __ldg(sa + i); dt[i] = __ldg(sb + i); }
    }
    __syncwarp();
    const uint32_t* sA = rows;
    const uint32_t* tA = rows + row_words;
    const uint32_t* sB = rows + (PK ? 2 : 0) * row_words;
    const uint32_t* tB = rows + (PK ? 3 : 1) * row_words;

    const int nmax = PK ? max(n[0], n[PAIRS - 1]) : n[0];
    const int steps = __reduce_max_sync(kFull, nmax) + G - 1;      // warp-uniform trip count
    const int max_col = row_words * 16 - 1;

    // ---- per-column state
    uint32_t up[T];        // G[i-1][j] for my columns (row 0: beta)
    uint32_t sel[T];       // PK: PRMT selector holding (tA[j], tB[j]);  else: t code
    const uint32_t beta2 = PK ? pack2(prm.beta) : (uint32_t)prm.beta;
#pragma unroll
    for (int c = 0; c < T; ++c) {
        int j = min(r * T + c, max_col);
        uint32_t ca = base_code(tA, j);
        if (PK) {
            uint32_t cb = base_code(tB, j);
            sel[c] = ca | 0x80u | ((4u + cb) << 8) | 0x8000u;
        } else {
            sel[c] = ca;
        }
        up[c] = beta2;
    }
    const uint32_t gu2 = PK ? pack2(prm.gu) : (uint32_t)prm.gu;
    const uint32_t gl2 = PK ? pack2(prm.gl) : (uint32_t)prm.gl;
    const uint32_t nbase2 = PK ? (uint32_t)(-prm.base) * 0x10001u : (uint32_t)(-prm.base);
    const uint32_t nev4 = (uint32_t)prm.nev * 0x01010101u;
    const uint32_t flip = (uint32_t)(prm.eqv ^ prm.nev);

    // running best of the last row, per half: value in G space, column j
    int bestv[PAIRS], bestj[PAIRS];
#pragma unroll
    for (int h = 0; h < PAIRS; ++h) {
        bestv[h] = (r == 0) ? prm.beta - n[h] * prm.base : INT_MIN;   // G[n][0]  (j = 0 floor, aligners.py:51-57)
        bestj[h] = 0;
    }

    uint32_t out = beta2;        // my last column of the row just finished (goes to lane r+1)
    uint32_t diag_in = beta2;    // G[i-1][rT-1] for the row about to be computed
    uint32_t col0 = beta2;       // lane 0: G[i][0] = beta - i*base

    for (int k = 0; k < steps; ++k) {
        const int i = k - r;                               // 0-based row of s handled this step
        uint32_t recv = __shfl_up_sync(kFull, out, 1, G);
        col0 += nbase2;                                    // lane 0 at step k: G[k+1][0]
        if (r == 0) recv = col0;
        if (i >= 0 && i < nmax) {
            uint32_t left = recv, diag = diag_in;
            if (PK) {
                uint32_t ca = base_code(sA, i), cb = base_code(sB, i);
                uint32_t lutA = nev4 ^ (flip << (8 * ca));
                uint32_t lutB = nev4 ^ (flip << (8 * cb));
#pragma unroll
                for (int c = 0; c < T; ++c) {
                    uint32_t sc = prmt(lutA, lutB, sel[c]);
                    uint32_t d = imad_add(sc, prm.one, diag);
                    uint32_t t1 = __viaddmax_s16x2(up[c], gu2, d);
                    uint32_t g = __viaddmax_s16x2(left, gl2, t1);
                    diag = up[c];
                    up[c] = g;
                    left = g;
                }
            } else {
                uint32_t ca = base_code(sA, i);
#pragma unroll
                for (int c = 0; c < T; ++c) {
                    int sc = (sel[c] == ca) ? prm.eqv : prm.nev;
                    int d = (int)diag + sc;
                    int t1 = __viaddmax_s32((int)up[c], (int)gu2, d);
                    int g = __viaddmax_s32((int)left, (int)gl2, t1);
                    diag = up[c];
                    up[c] = (uint32_t)g;
                    left = (uint32_t)g;
                }
            }
            out = left;
            diag_in = recv;
            // last row of a pair reached: fold my columns into its running (first) maximum
#pragma unroll
            for (int h = 0; h < PAIRS; ++h) {
                if (i + 1 == n[h]) {
#pragma unroll
                    for (int c = 0; c < T; ++c) {
                        int j = r * T + c + 1;
                        int v = PK ? (h == 0 ? half_lo(up[c]) : half_hi(up[c])) : (int)up[c];
                        if (j <= m[h] && v > bestv[h]) { bestv[h] = v; bestj[h] = j; }
                    }
                }
            }
        }
    }

    // ---- group arg-max: higher value wins, ties go to the smaller j (first maximum)
#pragma unroll
    for (int h = 0; h < PAIRS; ++h) {
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            int ov = __shfl_xor_sync(kFull, bestv[h], d, G);
            int oj = __shfl_xor_sync(kFull, bestj[h], d, G);
            if (ov > bestv[h] || (ov == bestv[h] && oj < bestj[h])) { bestv[h] = ov; bestj[h] = oj; }
        }
        int64_t p = p0 + h;
        if (r == 0 && p < P) {
            score_out[p] = bestv[h] - prm.beta + n[h] * prm.base;
            end_out[p] = bestj[h];
        }
    }
}

}  // namespace ovl
