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
// Staging.  The 2-bit packed rows of a warp's pairs (s and t of every pair, 48 B each at l = 150)
// are brought into shared memory by TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) that
// complete on one mbarrier per warp; everything after that reads shared memory only.
//
// Cost space.  The recurrence is evaluated on  C[i][j] = beta + i*maxs - H[i][j]  with
// maxs = max(match, mismatch):
//     diag:  C[i-1][j-1] + dc,   dc = maxs - score  in {0, |match - mismatch|}
//     up:    C[i-1][j]   + gu,   gu = maxs - indel
//     left:  C[i][j-1]   + gl,   gl = -indel
//     C[i][j] = min of the three;  C[0][j] = beta,  C[i][0] = beta + i*maxs
// For gap penalties (indel <= 0) every addend is NON-NEGATIVE and every C is in [0, Cmax], so
//   * two pairs are packed in the two 16-bit halves of each register (unsigned), and all three
//     adds are plain 32-bit adds with no carry between the halves -- they can be issued as
//     IMAD on the FMA pipe, next to the ALU pipe that does the DPX min (VIMNMX3.U16x2 /
//     VIADDMNMX.U16x2).  Measured on B200: ALU and FMA pipes each issue 64 lanes/clk/SM and
//     co-issue (probe kinds 5,6,7 in probe.cuh), so the per-column work is split over both.
//   * indel = -2^31 (the reference default: gapless) or any indel too negative to ever win is
//     replaced by Cmax+1, which is exact: such a candidate exceeds every true cell value.
// dc for both halves comes from ONE byte permute: the two query bases of a row select two
// 4-entry byte tables (lutA, lutB), precomputed per row in shared memory; the per-column
// selector register holds the two target bases (PRMT puts lutA[tA] in the low half,
// lutB[tB] in the high half; the zero bytes come from the sign-replicate selector applied to
// a byte < 128).
// Per column (2 cells) the kernel issues either
//     form 1:  PRMT, IMAD, VIADDMNMX, VIADDMNMX          (3 ALU + 1 FMA)
//     form 2:  PRMT, IMAD, IMAD, IMAD, VIMNMX3           (2 ALU + 3 FMA)
// mixed: OVL_DP_F2_NUM columns out of every OVL_DP_F2_DEN use form 2 (default 1 in 3, measured best).
//
// A 32-bit variant (one pair per group, s32 DPX ops, compare+select for dc) covers scoring
// schemes or read lengths whose range does not fit 16 bits.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "joinidx.cuh"

namespace ovl {

struct DpParams {
    int32_t eqc;        // maxs - match      (>= 0)   diagonal cost on equal bases
    int32_t nec;        // maxs - mismatch   (>= 0)   diagonal cost on different bases
    int32_t maxs;       // max(match, mismatch): per-row drift
    int32_t beta;       // bias so that every C >= 0
    int32_t gu;         // effective cost of an up move   (maxs - indel), clamped
    int32_t gl;         // effective cost of a left move  (-indel), clamped
    uint32_t one;       // == 1, a runtime value so the adds stay IMADs
    // the same constants pre-packed on the host (both 16-bit halves in packed mode); only read with
    // OVL_DP_PREPACK=1, which was measured slower than packing them in the kernel prologue
    uint32_t gu2, gl2, maxs2, beta2;
    int32_t cmax;       // upper bound of every cell value (cost space)
    int32_t gaps_never_win;   // both gap costs exceed cmax: no gap candidate can ever be the minimum
};

// When no gap can ever win (the reference's default indel = -2^31, overlapGraphs.py:53) the two gap costs
// are interchangeable with ANY constant above cmax.  The IMMG instantiation uses this fixed one, so that
// the gap adds carry an IMMEDIATE instead of a third register source: VIADDMNMX.U16x2 R, R, imm, R issues
// at twice the rate of the three-register form on this GPU (probe kinds 20 / 21).  Same recurrence,
// same instruction count, same results.  Requires cmax < kGapNever (so that it still never wins) and
// cmax + kGapNever <= 65535 (no carry between the halves): cmax <= 0x6fff.
#ifndef OVL_DP_GAPNEVER
#define OVL_DP_GAPNEVER 0x7000u
#endif
constexpr uint32_t kGapNever = OVL_DP_GAPNEVER;
constexpr uint32_t kGapNever2 = kGapNever * 0x10001u;

// optional fused edge expansion in the DP epilogue (all null: plain score/end output)
// OVL_DP_STREAM_HINTS=1 accesses the pair list (read once) and the edge rows (written once) with the streaming
// (evict-first) hints.  Measured: no effect on the kernel's DRAM traffic (35.0 GB read per launch at the 1 M-read
// workload either way, profiles/r2y_dp_traffic.csv), so it stays off.
#ifndef OVL_DP_STREAM_HINTS
#define OVL_DP_STREAM_HINTS 0
#endif
__device__ __forceinline__ int32_t ld_pair(const int32_t* p) {
#if OVL_DP_STREAM_HINTS
    return __ldcs(p);
#else
    return *p;
#endif
}
__device__ __forceinline__ void st_edge(int4* p, int4 v) {
#if OVL_DP_STREAM_HINTS
    __stcs(p, v);
#else
    *p = v;
#endif
}
struct DpEdgeOut {
    int4* edges;                 // int32[E][4] rows, or null
    const int32_t* copies;       // multiplicity per unique read, or null when every read occurs once
    const int64_t* node_off;     // exclusive scan of copies
    const int64_t* edge_off;     // explicit row offsets: exclusive scan of copies[a]*copies[b] over the pair list, or null
    JoinEdgeIndex join;          // implicit row offsets of a k-mer join (join.pair_off != null), see joinidx.cuh
};

// first edge row of pair p (local index) = (a, b) when reads have copies
__device__ __forceinline__ int4* dp_edge_dst(const DpEdgeOut& eo, int64_t p, int32_t a) {
    if (eo.join.pair_off != nullptr)
        return eo.edges + (join_edge_offset(eo.join, eo.copies, eo.join.p_begin + p, a) - eo.join.e_begin);
    return eo.edges + eo.edge_off[p];
}

#ifndef OVL_DP_THREADS
#define OVL_DP_THREADS 128
#endif
constexpr int kDpThreads = OVL_DP_THREADS;
#ifndef OVL_DP_MINB
#define OVL_DP_MINB 3          // resident CTAs per SM the register allocator must allow (168 regs; measured best)
#endif
#ifndef OVL_DP_MINB_IMM38
#define OVL_DP_MINB_IMM38 4    // the immediate-gap instantiation with 38 columns fits 127 registers without spilling: 4 CTAs per SM
                               // (measured 10.11 vs 9.84 TCUPS with 3; 32 columns x 32 lanes is best with 3, 5 spills)
#endif
#ifndef OVL_DP_F2_NUM
#define OVL_DP_F2_NUM 1        // columns using form 2 (FMA-heavy): NUM out of every DEN
#endif
#ifndef OVL_DP_PREPACK
#define OVL_DP_PREPACK 0       // 1: take the packed constants from the kernel parameters (measured slower: 8.4 vs
                               // 8.95 TCUPS; gap costs as immediates were slower too: 8.75 vs 9.09)
#endif
#ifndef OVL_DP_PLAIN_ADD
#define OVL_DP_PLAIN_ADD 0
#endif
#ifndef OVL_DP_KUNROLL
#define OVL_DP_KUNROLL 1       // unroll factor of the row loop
#endif
#ifndef OVL_DP_F2_DEN
#define OVL_DP_F2_DEN 3
#endif

constexpr int kDpKUnroll = OVL_DP_KUNROLL;
#ifndef OVL_DP_BULK
#define OVL_DP_BULK 1          // packed 2-bit instantiations of at most OVL_DP_BULK_MAX_COLS columns: prologue / epilogue on
#endif                         // aligned base windows and packed keys, row index kept in a register (dp_bulk() below)
#ifndef OVL_DP_PHASES
#define OVL_DP_PHASES 0        // 1: bulk instantiations split the row loop into ramp-up / steady / ramp-down phases
#endif
#ifndef OVL_DP_ROWVAR_ALL
#define OVL_DP_ROWVAR_ALL 1    // the register row index in the long instantiations too (32 x 32: 10.49 -> 10.58 TCUPS)
#endif
#ifndef OVL_DP_BULK_MAX_COLS
#define OVL_DP_BULK_MAX_COLS 256
#endif
// Which instantiations take the bulk prologue / epilogue.  Measured (kernel alone, profiles/r2v_dp_bulk_prologue.jsonl):
// 4 x 38 at l = 150: 10.11 -> 10.66 TCUPS (the prologue and epilogue were 7 % of the executed instructions there);
// 32 x 32 at l = 1,000, where they are amortised over 1,031 rows: 10.48 -> 10.35, so the long instantiations keep the
// row-strided code.
template <int G, int T, bool PK, int BITS>
__host__ __device__ constexpr bool dp_bulk() { return OVL_DP_BULK && PK && BITS == 2 && G * T <= OVL_DP_BULK_MAX_COLS; }

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// a + c as IMAD (a * 1 + c): runs on the FMA pipe instead of the ALU pipe
__device__ __forceinline__ uint32_t fma_add(uint32_t a, uint32_t one, uint32_t c) {
#if OVL_DP_PLAIN_ADD
    (void)one;
    return a + c;                  // let ptxas pick IADD3 (ALU) or IMAD.IADD with an immediate 1 (FMA)
#else
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
    return d;
#endif
}
__device__ __forceinline__ uint32_t base_code(const uint32_t* row, int i) {
    return (row[i >> 4] >> ((i & 15) * 2)) & 3u;
}
// T consecutive 2-bit bases of a packed row, starting at base `first`, re-aligned so that base c of the window sits
// at bits 2 (c % 16) of w[c / 16]: one funnel shift per word; after that every per-base extraction shifts by a
// compile-time amount.  Words past the end of the row are read from its last word (their bases only ever feed
// padded DP cells, which nothing reads).
template <int T>
__device__ __forceinline__ void base_window(const uint32_t* row, int row_words, int first, uint32_t (&w)[(2 * T + 31) / 32]) {
    constexpr int NW = (2 * T + 31) / 32;
    const int i0 = first >> 4;
    const int sh = (first & 15) * 2;
    uint32_t a[NW + 1];
    OVL_CHECK(first >= 0 && row_words >= 1);
#pragma unroll
    for (int i = 0; i <= NW; ++i) a[i] = row[min(i0 + i, row_words - 1)];
#pragma unroll
    for (int i = 0; i < NW; ++i) w[i] = __funnelshift_r(a[i], a[i + 1], sh);
}
// ((w >> s) & 3) << up, with compile-time s and up
__device__ __forceinline__ uint32_t base_field(uint32_t w, int s, int up) {
    const uint32_t mask = 3u << up;
    return s >= up ? (w >> (s - up)) & mask : (w << (up - s)) & mask;
}
// BITS = 2: 2-bit packed rows (16 bases per word); BITS = 8: byte rows (any alphabet)
template <int BITS>
__device__ __forceinline__ uint32_t read_symbol(const uint32_t* row, int i) {
    return BITS == 2 ? base_code(row, i) : (uint32_t)reinterpret_cast<const uint8_t*>(row)[i];
}
__device__ __forceinline__ uint32_t pack2(int v) { return ((uint32_t)v & 0xffffu) * 0x10001u; }
__device__ __forceinline__ constexpr bool dp_form2(int c) {
    return OVL_DP_F2_NUM > 0 && (c * OVL_DP_F2_NUM) % OVL_DP_F2_DEN < OVL_DP_F2_NUM;
}
// Column form by position.  With OVL_DP_PATTERN (decimal digits, most significant first, repeated over the columns):
//   1  PRMT, add, VIADDMNMX(up), VIADDMNMX(left)
//   2  PRMT, 3 adds, VIMNMX3
//   3  PRMT, add, up+g on the FMA pipe, two-input packed min, VIADDMNMX(left)
//   4  as 3 with the up+g add left to ptxas
//   5  PRMT, add, VIADDMNMX(up), left+g on the FMA pipe, two-input packed min
//   6  PRMT, add, VIADDMNMX(up), left+g, fp16x2 min      (6-8 need every value < 0x7c00: OVL_DP_GAPNEVER=0x3e00)
//   7  PRMT, 3 adds, two fp16x2 mins
//   8  PRMT, add, up+g, fp16x2 min, VIADDMNMX(left)
// The two-input packed min (VIMNMX3.U16x2 with a repeated source) issues at twice the rate of the
// three-source DPX forms (probe kind 17).
#ifdef OVL_DP_PATTERN
__device__ __forceinline__ constexpr int dp_form(int c) {
    int nd = 0;
    for (long long v = OVL_DP_PATTERN; v > 0; v /= 10) ++nd;
    long long v = OVL_DP_PATTERN;
    for (int skip = nd - 1 - c % nd; skip > 0; --skip) v /= 10;
    return (int)(v % 10);
}
#else
__device__ __forceinline__ constexpr int dp_form(int c) { return dp_form2(c) ? 2 : 1; }
#endif
// the short (bulk) instantiations run 2 form-2 columns in 5 (10.72 vs 10.66 TCUPS at 4 x 38 with 1 in 3); the long ones 1 in 3
template <bool BULK_>
__device__ __forceinline__ constexpr int dp_form_of(int c) {
#if defined(OVL_DP_PATTERN) || defined(OVL_DP_F2_FIXED)
    return dp_form(c);
#else
    return BULK_ ? ((c * 2) % 5 < 2 ? 2 : 1) : dp_form(c);
#endif
}
// min of two packed halves that are both below 0x7c00, as an fp16x2 min: the bit patterns of non-negative finite
// halves order like the integers they spell (no .ftz: subnormal patterns are kept), SASS HMNMX2
__device__ __forceinline__ uint32_t hmin2_bits(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// a + c on the FMA pipe in every instantiation (a * one + c, `one` a runtime 1)
__device__ __forceinline__ uint32_t fma_add_always(uint32_t a, uint32_t one, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
    return d;
}
// (There is no cheaper 2-input packed min to build a third form from: __vminu2 compiles to
// VIMNMX3.U16x2 with a repeated operand on sm_100a -- probe kind 17.)

// ---- TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
#ifndef OVL_DP_TMA_EVICT_LAST
#define OVL_DP_TMA_EVICT_LAST 0   // 1: the bulk copies of packed read rows carry an L2 evict_last policy (the rows are re-read by every pair)
#endif
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
#if OVL_DP_TMA_EVICT_LAST
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar), "l"(pol) : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
#endif
}

// rows of the per-couple LUT array, padded so that consecutive couples start 8 banks apart
__host__ __device__ inline int dp_lut_rows(int max_len) {
    int r = max_len < 1 ? 1 : max_len;
    return ((r + 11) / 16) * 16 + 4;       // == 4 (mod 16), >= r
}

// PK = true : two pairs per group (unsigned 16-bit halves).  PK = false: one pair per group (s32).
// BITS = 8 (byte-coded reads, any alphabet; packed mode only): the diagonal cost comes from
// XOR + min(.,1) + multiply-add (LOP3, VIMNMX.U16x2, IMAD) instead of the 4-entry PRMT table;
// requires eqc == 0 (match >= mismatch).
template <int G, int T, bool PK, int BITS = 2, bool IMMG = false>
__global__ void __launch_bounds__(kDpThreads, (T > 38 ? 2 : (IMMG && T == 38) ? OVL_DP_MINB_IMM38 : OVL_DP_MINB)) overlap_dp_kernel(
    const uint32_t* __restrict__ packed, int row_words, const int32_t* __restrict__ len,
    const int32_t* __restrict__ pair_a, const int32_t* __restrict__ pair_b, int64_t P, int lut_rows,
    DpParams prm, int32_t* __restrict__ score_out, int32_t* __restrict__ end_out, DpEdgeOut eo) {
    constexpr int PAIRS = PK ? 2 : 1;
    constexpr int GROUPS_PER_WARP = 32 / G;
    constexpr int GROUPS_PER_CTA = (kDpThreads / 32) * GROUPS_PER_WARP;
    constexpr int ROWS_PER_GROUP = 2 * PAIRS;          // s and t of each pair
    // dynamic shared memory: [packed read rows][per-row score tables][one mbarrier per warp]
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* smem_rows = reinterpret_cast<uint32_t*>(smem_raw);              // [GROUPS_PER_CTA][ROWS_PER_GROUP][row_words]
    uint2* smem_lut = reinterpret_cast<uint2*>(smem_rows + (size_t)GROUPS_PER_CTA * ROWS_PER_GROUP * row_words);
    uint64_t* smem_bar = reinterpret_cast<uint64_t*>(smem_lut + (size_t)GROUPS_PER_CTA * lut_rows);

    const unsigned lane = lane_id();
    const int r = lane % G;                            // lane within the group
    const int gib = (threadIdx.x >> 5) * GROUPS_PER_WARP + lane / G;   // group within the CTA
    const int64_t grp = (int64_t)blockIdx.x * GROUPS_PER_CTA + gib;
    const int64_t p0 = grp * PAIRS;

    // ---- stage the packed reads of this warp's pairs in shared memory with TMA bulk copies:
    // one 1-D copy per read row (row_words*4 bytes, 16-byte aligned at both ends), all completing
    // on the warp's mbarrier.  Lane l issues the copies of rows l, l+32, ... of the warp.
    const int wib = threadIdx.x >> 5;
    const uint32_t bar = smem_u32(smem_bar + wib);
    uint32_t* warp_rows = smem_rows + (size_t)wib * GROUPS_PER_WARP * ROWS_PER_GROUP * row_words;
    {
        constexpr int NROWS = GROUPS_PER_WARP * ROWS_PER_GROUP;
        const uint32_t row_bytes = (uint32_t)row_words * 4u;
        if (lane == 0) {
            mbar_init(bar, 1);
            mbar_expect_tx(bar, NROWS * row_bytes);
        }
        __syncwarp();
        const int64_t warp_p0 = ((int64_t)blockIdx.x * GROUPS_PER_CTA + (int64_t)wib * GROUPS_PER_WARP) * PAIRS;
        for (int idx = lane; idx < NROWS; idx += 32) {
            int64_t p = warp_p0 + idx / 2;                  // rows come in (s, t) order per pair
            int32_t uid = 0;
            if (p < P) uid = (idx & 1) ? ld_pair(pair_b + p) : ld_pair(pair_a + p);
            tma_load_1d(smem_u32(warp_rows + (size_t)idx * row_words), packed + (size_t)uid * row_words, row_bytes, bar);
        }
    }

    int32_t n[PAIRS], m[PAIRS];
    const uint32_t* srow[PAIRS];
    const uint32_t* trow[PAIRS];
    uint32_t* grp_rows = smem_rows + (size_t)gib * ROWS_PER_GROUP * row_words;
#pragma unroll
    for (int h = 0; h < PAIRS; ++h) {
        int64_t p = p0 + h;
        bool live = p < P;
        int32_t a = live ? ld_pair(pair_a + p) : 0, b = live ? ld_pair(pair_b + p) : 0;
        n[h] = live ? len[a] : 0;
        m[h] = live ? len[b] : 0;
        srow[h] = grp_rows + (size_t)(2 * h) * row_words;
        trow[h] = grp_rows + (size_t)(2 * h + 1) * row_words;
    }
    // the rows have landed (phase 0 completes once); a copy that never completes (bad pointer) traps
    // instead of hanging the GPU
    for (uint32_t spins = 0; !mbar_try_wait(bar, 0); )
        if (++spins > (1u << 20)) __trap();
    const int nmax = PK ? max(n[0], n[PAIRS - 1]) : n[0];
    const int max_col = row_words * (BITS == 2 ? 16 : 4) - 1;

    // ---- per-row tables: lut.x / lut.y = bytes {cost of s[i] vs code 0..3} for pair 0 / 1
    uint2* lut = smem_lut + (size_t)gib * lut_rows;
    constexpr bool BULK = dp_bulk<G, T, PK, BITS>();
    constexpr int NW = (2 * T + 31) / 32;
    if (BULK) {
        // lane r builds rows [rT, rT + T): the two query reads' bases come out of two aligned windows with
        // compile-time shifts (the row-strided version spent ~20 instructions per row on variable shifts)
        const uint32_t nec4 = (uint32_t)prm.nec * 0x01010101u;
        const uint32_t flip = (uint32_t)(prm.eqc ^ prm.nec);
        uint32_t sa[NW], sb[NW];
        base_window<T>(srow[0], row_words, r * T, sa);
        base_window<T>(srow[PAIRS - 1], row_words, r * T, sb);
#pragma unroll
        for (int c = 0; c < T; ++c) {
            const int i = r * T + c;
            const int sft = 2 * (c & 15);
            const uint32_t xa = nec4 ^ (flip << base_field(sa[c >> 4], sft, 3));      // byte ca of the table = eqc
            const uint32_t xb = nec4 ^ (flip << base_field(sb[c >> 4], sft, 3));
            OVL_CHECK(i >= 0 && i < lut_rows);
            lut[i] = make_uint2(xa, xb);         // unconditional: the table has G*T rows (launch_dp); rows >= nmax are never read
        }
    } else {
        const uint32_t nec4 = (uint32_t)prm.nec * 0x01010101u;
        const uint32_t flip = (uint32_t)(prm.eqc ^ prm.nec);
        for (int i = r; i < nmax; i += G) {
            uint32_t ca = read_symbol<BITS>(srow[0], min(i, max_col));
            uint32_t cb = read_symbol<BITS>(srow[PAIRS - 1], min(i, max_col));
            OVL_CHECK(i < lut_rows);
            if (BITS == 8)      lut[i] = make_uint2(PK ? (ca | (cb << 16)) : ca, 0u);     // the two query symbols
            else                lut[i] = PK ? make_uint2(nec4 ^ (flip << (8 * ca)), nec4 ^ (flip << (8 * cb))) : make_uint2(ca, 0u);
        }
    }
    __syncwarp();
    const uint32_t lut_addr = (uint32_t)__cvta_generic_to_shared(lut);

    // ---- per-column state
    uint32_t up[T];        // C[i-1][j] for my columns (row 0: beta)
    uint32_t sel[T];       // PK: PRMT selector holding (tA[j], tB[j]);  else: t code
#if OVL_DP_PREPACK
    const uint32_t beta2 = prm.beta2;
#else
    const uint32_t beta2 = PK ? pack2(prm.beta) : (uint32_t)prm.beta;
#endif
    if (BULK) {
        uint32_t wa[NW], wb[NW];
        base_window<T>(trow[0], row_words, r * T, wa);
        base_window<T>(trow[PAIRS - 1], row_words, r * T, wb);
#pragma unroll
        for (int c = 0; c < T; ++c) {
            const int sft = 2 * (c & 15);
            // ca | 0x80 | ((4 + cb) << 8) | 0x8000
            sel[c] = base_field(wa[c >> 4], sft, 0) | base_field(wb[c >> 4], sft, 8) | 0x8480u;
            up[c] = beta2;
        }
    } else {
#pragma unroll
        for (int c = 0; c < T; ++c) {
            int j = min(r * T + c, max_col);
            uint32_t ca = read_symbol<BITS>(trow[0], j);
            if (PK) {
                uint32_t cb = read_symbol<BITS>(trow[PAIRS - 1], j);
                sel[c] = BITS == 8 ? (ca | (cb << 16)) : (ca | 0x80u | ((4u + cb) << 8) | 0x8000u);
            } else {
                sel[c] = ca;
            }
            up[c] = beta2;
        }
    }
#if OVL_DP_PREPACK
    const uint32_t gu2 = prm.gu2, gl2 = prm.gl2, maxs2 = prm.maxs2;
#else
    const uint32_t gu2 = IMMG ? kGapNever2 : PK ? pack2(prm.gu) : (uint32_t)prm.gu;
    const uint32_t gl2 = IMMG ? kGapNever2 : PK ? pack2(prm.gl) : (uint32_t)prm.gl;
    const uint32_t maxs2 = PK ? (uint32_t)prm.maxs * 0x10001u : (uint32_t)prm.maxs;
#endif
    const uint32_t one = prm.one;

    // running best of the last row, per pair: cost (smaller is better), column j
    int bestv[PAIRS], bestj[PAIRS];
#pragma unroll
    for (int h = 0; h < PAIRS; ++h) {
        bestv[h] = (r == 0) ? prm.beta + n[h] * prm.maxs : INT_MAX;   // C[n][0]  (j = 0, aligners.py:51-57)
        bestj[h] = 0;
    }

    // the shorter pair of the couple (if any), folded inside the loop
    const int hshort = (PK && n[PAIRS - 1] < n[0]) ? PAIRS - 1 : 0;
    const int nshort = n[hshort], mshort = m[hshort];
    int bests = bestv[hshort], bestjs = 0;

    uint32_t out = beta2;        // my last column of the row just finished (goes to lane r+1)
    uint32_t diag_in = beta2;    // C[i-1][rT-1] for the row about to be computed
    uint32_t col0 = beta2;       // lane 0: C[i][0] = beta + i*maxs
    const int steps = __reduce_max_sync(kFull, nmax) + G - 1;      // warp-uniform trip count

    // BULK: the lane's row index lives in a register (one add per step) instead of being re-derived from the step
    // counter and the thread index (S2R, LOP3, IADD3 per step under the 128-register cap)
    constexpr bool ROWV = BULK || OVL_DP_ROWVAR_ALL;
    int irow = -r;
    const int fold_row = (PK && nshort < nmax) ? nshort - 1 : INT_MIN;     // the step at which the shorter pair ends
    // One wavefront step.  CHECKED: the lane may be outside its rows (ramp-up / ramp-down of the wavefront) and the
    // shorter pair of the couple may end at this row.  Unchecked steps are the steady state, see the phases below.
    auto step = [&](int k, auto checked_tag) {
        constexpr bool CHECKED = decltype(checked_tag)::value;
        const int i = ROWV ? irow++ : k - r;               // 0-based row of s handled this step
        uint32_t recv = __shfl_up_sync(kFull, out, 1, G);
        col0 += maxs2;                                     // lane 0 at step k: C[k+1][0]
        if (r == 0) recv = col0;
        if (!CHECKED || (unsigned)i < (unsigned)nmax) {   // 0 <= i < nmax in one compare
            uint32_t left = recv, diag = diag_in;
            uint2 lu;
            OVL_CHECK(i >= 0 && i < lut_rows);
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lu.x), "=r"(lu.y) : "r"(lut_addr + 8u * (unsigned)i));
#pragma unroll
            for (int c = 0; c < T; ++c) {
                uint32_t g;
                if (PK) {
                    uint32_t a1;
                    if (BITS == 8) {
                        uint32_t ne = __vminu2(sel[c] ^ lu.x, 0x00010001u);      // 1 per half where the symbols differ
                        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a1) : "r"(ne), "r"((uint32_t)prm.nec), "r"(diag));
                    } else {
                        uint32_t dc = prmt(lu.x, lu.y, sel[c]);
                        a1 = IMMG ? dc + diag : fma_add(dc, one, diag);     // IMMG: two register sources either way
                    }
                    if (dp_form_of<BULK>(c) == 2) {
                        // IMMG: x + immediate (ptxas emits VIADD).  Forcing these two adds onto the FMA pipe as
                        // IMAD x, one, imm was measured SLOWER (9.80 vs 10.11 TCUPS), at every form mix.
                        uint32_t a2 = IMMG ? up[c] + gu2 : fma_add(up[c], one, gu2);
                        uint32_t a3 = IMMG ? left + gl2 : fma_add(left, one, gl2);
                        g = __vimin3_u16x2(a1, a2, a3);
                    } else if (dp_form_of<BULK>(c) == 3 || dp_form_of<BULK>(c) == 4) {
                        uint32_t a2 = dp_form_of<BULK>(c) == 3 ? fma_add_always(up[c], one, gu2) : up[c] + gu2;
                        uint32_t m1 = __vminu2(a1, a2);
                        g = __viaddmin_u16x2(left, gl2, m1);
                    } else if (dp_form_of<BULK>(c) == 6) {
                        uint32_t t1 = __viaddmin_u16x2(up[c], gu2, a1);
                        g = hmin2_bits(t1, left + gl2);
                    } else if (dp_form_of<BULK>(c) == 7) {
                        g = hmin2_bits(hmin2_bits(a1, up[c] + gu2), left + gl2);
                    } else if (dp_form_of<BULK>(c) == 8) {
                        uint32_t m1 = hmin2_bits(a1, up[c] + gu2);
                        g = __viaddmin_u16x2(left, gl2, m1);
                    } else if (dp_form_of<BULK>(c) == 5) {
                        uint32_t t1 = __viaddmin_u16x2(up[c], gu2, a1);
                        uint32_t a3 = fma_add_always(left, one, gl2);
                        g = __vminu2(t1, a3);
                    } else {
                        uint32_t t1 = __viaddmin_u16x2(up[c], gu2, a1);
                        g = __viaddmin_u16x2(left, gl2, t1);
                    }
                } else {
                    int dc = (sel[c] == lu.x) ? prm.eqc : prm.nec;
                    int a1 = (int)diag + dc;
                    int t1 = __viaddmin_s32((int)up[c], (int)gu2, a1);
                    g = (uint32_t)__viaddmin_s32((int)left, (int)gl2, t1);
                }
                diag = up[c];
                up[c] = g;
                left = g;
            }
            out = left;
            diag_in = recv;
            // a pair shorter than its partner reaches its last row inside the loop: fold my columns
            // into its running (first) minimum now, its half keeps computing an ignored padded DP
            if (CHECKED && PK && (ROWV ? i == fold_row : (i + 1 == nshort && nshort < nmax))) {
#pragma unroll
                for (int c = 0; c < T; ++c) {
                    int j = r * T + c + 1;
                    int v = (int)(hshort == 0 ? (up[c] & 0xffffu) : (up[c] >> 16));
                    if (j <= mshort && v < bests) { bests = v; bestjs = j; }
                }
            }
        }
    };
    if (BULK && OVL_DP_PHASES) {
        // Three phases with warp-uniform bounds.  For G - 1 <= k < (shortest read of the warp) - 1 every lane is inside
        // its rows (0 <= k - r < n) and no pair ends (that needs k - r == n - 1, k >= n - 1), so the steady state runs
        // without the range check, the fold check and the reconvergence bracket around them: 3 ALU instructions and 4
        // control instructions fewer per step.  The few steps before and after take the checked path.
        int k = 0;
        const int k_lo = min(G - 1, steps);
        for (; k < k_lo; ++k) step(k, std::true_type{});
        const int k_hi = max(k_lo, min(steps, __reduce_min_sync(kFull, nshort) - 1));
        for (; k < k_hi; ++k) step(k, std::false_type{});
        for (; k < steps; ++k) step(k, std::true_type{});
    } else {
#pragma unroll kDpKUnroll
        for (int k = 0; k < steps; ++k) step(k, std::true_type{});
    }
    // the pair(s) with n == nmax: after the loop every lane still holds its columns of the last row
#pragma unroll
    for (int h = 0; h < PAIRS; ++h) {
        if (n[h] == nmax && nmax > 0) {
            if (BULK) {
                // first minimum over my columns inside t as ONE running min of (value << 8 | column): the smallest
                // value wins, then the smallest column; then against what I hold (lane 0: the j = 0 cell)
                const int nv = m[h] - r * T;                 // columns c < nv have j = rT + c + 1 <= m
                int key = INT_MAX;
#pragma unroll
                for (int c = 0; c < T; ++c) {
                    const int kc = (int)(h == 0 ? (up[c] & 0xffffu) : (up[c] >> 16)) * 256 + c;
                    if (c < nv) key = min(key, kc);
                }
                if (key != INT_MAX && (key >> 8) < bestv[h]) { bestv[h] = key >> 8; bestj[h] = r * T + (key & 0xff) + 1; }
            } else {
#pragma unroll
                for (int c = 0; c < T; ++c) {
                    int j = r * T + c + 1;
                    int v = PK ? (int)(h == 0 ? (up[c] & 0xffffu) : (up[c] >> 16)) : (int)up[c];
                    if (j <= m[h] && v < bestv[h]) { bestv[h] = v; bestj[h] = j; }
                }
            }
        } else if (PK) {
            bestv[h] = bests;
            bestj[h] = bestjs;
        }
    }

    // ---- group arg-min: lower cost wins, ties go to the smaller j (first maximum of H)
#pragma unroll
    for (int h = 0; h < PAIRS; ++h) {
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            int ov = __shfl_xor_sync(kFull, bestv[h], d, G);
            int oj = __shfl_xor_sync(kFull, bestj[h], d, G);
            if (ov < bestv[h] || (ov == bestv[h] && oj < bestj[h])) { bestv[h] = ov; bestj[h] = oj; }
        }
        int64_t p = p0 + h;
        if (r == 0 && p < P) {
            const int32_t score = prm.beta + n[h] * prm.maxs - bestv[h];
            OVL_CHECK(p >= 0 && p < P);
            if (eo.edges == nullptr) {
                score_out[p] = score;
                end_out[p] = bestj[h];
            } else {
                // fused K6 (overlapGraphs.py:55-60): emit the pair's copy_a x copy_b edge rows directly
                const int32_t a = ld_pair(pair_a + p), b = ld_pair(pair_b + p);
                if (eo.copies == nullptr) {
                    st_edge(eo.edges + p, make_int4(a, b, score, bestj[h]));
                } else {
                    const int32_t ca = eo.copies[a], cb = eo.copies[b];
                    const int32_t na = (int32_t)eo.node_off[a], nb = (int32_t)eo.node_off[b];
                    int4* dst = dp_edge_dst(eo, p, a);
                    for (int32_t ia = 0; ia < ca; ++ia)
                        for (int32_t ib = 0; ib < cb; ++ib) st_edge(dst++, make_int4(na + ia, nb + ib, score, bestj[h]));
                }
            }
        }
    }
}

// ------------------------------------------------------------------ long reads
// Reads longer than the register wavefront covers (32 lanes x 76 columns): one CTA per pair sweeps
// the anti-diagonals with three rolling diagonals in shared memory (int32 cost space, same
// recurrence, same first-minimum rule).  Slower per cell than the packed kernel, any length whose
// diagonals fit shared memory (3 * (n+1) ints).
constexpr int kDpLongThreads = 256;

template <int BITS>
__global__ void __launch_bounds__(kDpLongThreads) overlap_dp_long_kernel(
    const uint32_t* __restrict__ packed, int row_words, const int32_t* __restrict__ len,
    const int32_t* __restrict__ pair_a, const int32_t* __restrict__ pair_b, int64_t P,
    DpParams prm, int32_t* __restrict__ score_out, int32_t* __restrict__ end_out, DpEdgeOut eo) {
    extern __shared__ int32_t diag_l[];
    __shared__ int s_best[kDpLongThreads / 32], s_bj[kDpLongThreads / 32];
    const int64_t p = blockIdx.x;
    if (p >= P) return;
    const int32_t a = pair_a[p], b = pair_b[p];
    const int n = len[a], m = len[b];
    const uint32_t* srow = packed + (size_t)a * row_words;
    const uint32_t* trow = packed + (size_t)b * row_words;
    const int stride = n + 1;
    // C[i][j] = beta + i*maxs - H[i][j];  row 0: beta, column 0: beta + i*maxs
    for (int i = threadIdx.x; i < 3 * stride; i += blockDim.x) diag_l[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        diag_l[0] = prm.beta;                                   // d = 0: C[0][0]
        diag_l[stride + 0] = prm.beta;                          // d = 1: C[0][1]
        if (n >= 1) diag_l[stride + 1] = prm.beta + prm.maxs;   //        C[1][0]
    }
    int bestv = INT_MAX, bestj = 0;                             // last row: first minimum over j >= 1
    __syncthreads();
    for (int d = 2; d <= n + m; ++d) {
        int32_t* cur = diag_l + (d % 3) * stride;
        const int32_t* p1 = diag_l + ((d - 1) % 3) * stride;
        const int32_t* p2 = diag_l + ((d - 2) % 3) * stride;
        int ilo = max(1, d - m), ihi = min(n, d - 1);
        for (int i = ilo + (int)threadIdx.x; i <= ihi; i += blockDim.x) {
            int j = d - i;
            int dc = read_symbol<BITS>(srow, i - 1) == read_symbol<BITS>(trow, j - 1) ? prm.eqc : prm.nec;
            int v = min(min(p2[i - 1] + dc, p1[i - 1] + prm.gu), p1[i] + prm.gl);
            cur[i] = v;
            if (i == n && v < bestv) { bestv = v; bestj = j; }   // one thread owns row n: j ascending
        }
        if (threadIdx.x == 0) {
            if (d <= m) cur[0] = prm.beta;                       // C[0][d]
            if (d <= n) cur[d] = prm.beta + d * prm.maxs;        // C[d][0]
        }
        __syncthreads();
    }
    // the cells of row n are visited by different threads (i == n lands on thread (n - ilo) % blockDim):
    // reduce (value, then smaller j)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        int ov = __shfl_xor_sync(kFull, bestv, off), oj = __shfl_xor_sync(kFull, bestj, off);
        if (ov < bestv || (ov == bestv && oj < bestj)) { bestv = ov; bestj = oj; }
    }
    if (lane_id() == 0) { s_best[threadIdx.x >> 5] = bestv; s_bj[threadIdx.x >> 5] = bestj; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kDpLongThreads / 32; ++w)
            if (s_best[w] < bestv || (s_best[w] == bestv && s_bj[w] < bestj)) { bestv = s_best[w]; bestj = s_bj[w]; }
        int c0 = prm.beta + n * prm.maxs;                        // j = 0 (aligners.py:51-57): wins ties, it comes first
        if (n == 0 || m == 0 || c0 <= bestv) { bestv = c0; bestj = 0; }
        const int32_t score = prm.beta + n * prm.maxs - bestv;
        if (eo.edges == nullptr) {
            score_out[p] = score;
            end_out[p] = bestj;
        } else if (eo.copies == nullptr) {
            eo.edges[p] = make_int4(a, b, score, bestj);
        } else {
            const int32_t ca = eo.copies[a], cb = eo.copies[b];
            const int32_t na = (int32_t)eo.node_off[a], nb = (int32_t)eo.node_off[b];
            int4* dst = dp_edge_dst(eo, p, a);
            for (int32_t ia = 0; ia < ca; ++ia)
                for (int32_t ib = 0; ib < cb; ++ib) *dst++ = make_int4(na + ia, nb + ib, score, bestj);
        }
    }
}

}  // namespace ovl
