// K2 / K3: the prefix index and the candidate join (replaces overlapGraphs.py:30-52).
//
//   K2  stable LSD radix sort of (prefix_key, uid) on key_bits bits with D-bit digits (D <= 10):
//       key_bits <= 10 (k <= 5) sorts in ONE pass, k <= 10 in two.  Stability keeps uids ascending
//       inside a bucket = the reference's bucket-append order (overlapGraphs.py:38-40).
//       On top of the sorted keys sits a direct-address bucket table over the key's top
//       table_bits bits: table[t] = first sorted position whose key >> shift is >= t.  When the whole
//       key fits (4^k <= 2^22) a bucket is table[key] .. table[key+1] -- no search at all; longer
//       keys finish with a binary search inside that (tiny) range.
//   K3  join_count: per source read a, bucket of suffix_key[a], a's own slot in it (pos_of, written
//       by the last sort pass), candidate count and -- with duplicate reads -- edge count;
//       one (pair) scan; join_finalize: totals + this rank's slice + D2H chunk bounds, written
//       where the host can read them without a further kernel; join_fill: the (a, b) list.
#pragma once
#include "common.cuh"
#include "joinidx.cuh"
#include "scan.cuh"

namespace ovl {

constexpr int kSortWarps = 8;            // warps per CTA
constexpr int kSortThreads = kSortWarps * 32;
#ifndef OVL_SORT_ROUNDS
#define OVL_SORT_ROUNDS 16
#endif
constexpr int kSortRounds = OVL_SORT_ROUNDS;           // elements per lane
constexpr int kSortWarpChunk = 32 * kSortRounds;             // 512 consecutive elements owned by one warp
constexpr int kSortChunk = kSortWarps * kSortWarpChunk;      // 4,096 consecutive elements owned by one CTA (8 rounds / 2,048: 615 -> 550 us for three passes at 8 M reads, 105 -> 94 us for two at 1 M)
constexpr size_t kSortStageBytes = (size_t)kSortChunk * (sizeof(uint64_t) + sizeof(uint32_t));    // dynamic shared memory of the scatter kernel
constexpr int kSortMaxDigit = 10;        // bits per pass: 8 warps x 1024 counters = 32 KB of shared memory
constexpr int kTableMaxBits = 22;        // direct-address table: at most 4 Mi + 1 entries (16 MB)

__host__ __device__ inline int sort_passes(int key_bits) { return (key_bits + kSortMaxDigit - 1) / kSortMaxDigit; }
__host__ __device__ inline int sort_digit_bits(int key_bits) {
    int p = sort_passes(key_bits);
    return (key_bits + p - 1) / p;
}

// per-CTA digit histogram of the CTA's chunk -> hist[digit * C + cta]   (C = number of CTAs)
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const uint64_t* __restrict__ keys,
                                                                 const int32_t* __restrict__ len, int k,
                                                                 const int64_t* __restrict__ n_ptr, int64_t n_static,
                                                                 int shift, int digit_bits, int64_t C,
                                                                 int32_t* __restrict__ hist) {
    __shared__ int32_t cnt[1 << kSortMaxDigit];
    const int nd = 1 << digit_bits;
    const unsigned mask = (unsigned)nd - 1u;
    for (int i = threadIdx.x; i < nd; i += kSortThreads) cnt[i] = 0;
    __syncthreads();
    const int64_t n = FIRST ? n_static : *n_ptr;
    const int64_t base = (int64_t)blockIdx.x * kSortChunk;
    uint64_t key[kSortRounds];
    bool live[kSortRounds];
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {                       // all loads before the first atomic
        int64_t idx = base + r * kSortThreads + threadIdx.x;
        live[r] = idx < n;
        key[r] = live[r] ? keys[idx] : 0;
        if (FIRST && live[r]) live[r] = len[idx] >= k;
    }
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r)
        if (live[r]) atomicAdd(&cnt[(unsigned)(key[r] >> shift) & mask], 1);
    __syncthreads();
    for (int d = threadIdx.x; d < nd; d += kSortThreads) hist[(int64_t)d * C + blockIdx.x] = cnt[d];
}

// One CTA per digit: in-place exclusive scan of the digit's per-CTA counts hist[d * C .. d * C + C) and the digit's
// total.  Together with the scan of the (at most 1,024) digit totals that every scatter CTA does for itself this
// replaces the three-kernel scan over the whole [digit][cta] array: one launch between histogram and scatter.
__global__ void __launch_bounds__(kScanThreads) sort_digit_scan_kernel(int32_t* __restrict__ hist, int64_t C,
                                                                       int32_t* __restrict__ digit_total) {
    __shared__ int32_t sm[kScanThreads / 32 + 1];
    __shared__ int32_t tile[scan_smem_elems<int32_t>()];
    int32_t* mine = hist + (int64_t)blockIdx.x * C;
    const LoadArray<int32_t> f{mine};
    const StoreArray<int32_t> store{mine};
    int32_t carry = 0;
    for (int64_t start = 0; start < C; start += scan_tile<int32_t>())
        carry += scan_one_tile<LoadArray<int32_t>, StoreArray<int32_t>, int32_t>(f, store, start, C, carry, tile, sm);
    if (threadIdx.x == 0) digit_total[blockIdx.x] = carry;
}

// Stable scatter of the CTA's chunk to the scanned offsets.  Warp w owns elements [256 w, 256 w + 256) of the
// chunk and ranks them among themselves (8 rounds of __match_any_sync); the CTA scans the per-warp digit
// counts into CTA-local sorted positions, the elements are staged in shared memory in that order, and
// consecutive threads then write consecutive elements: every run of equal digits goes out as one
// contiguous, coalesced piece.
// FIRST: input is (prefix_key, uid = index) straight from K1 and reads shorter than k are dropped; n_out
// receives the number of survivors.  Last pass only: pos_of[uid] = sorted position, and -- when reads have
// copies -- sorted_copies[position] = copies[uid] (what the join scans into `cum`).
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const uint64_t* __restrict__ keys_in,
                                                                    const uint32_t* __restrict__ uid_in,
                                                                    const int32_t* __restrict__ len, int k,
                                                                    const int64_t* __restrict__ n_ptr, int64_t n_static,
                                                                    int shift, int digit_bits, int64_t C,
                                                                    const int32_t* __restrict__ hist_scanned,
                                                                    const int32_t* __restrict__ digit_total,
                                                                    int32_t* __restrict__ table_out,
                                                                    uint64_t* __restrict__ keys_out,
                                                                    uint32_t* __restrict__ uid_out,
                                                                    int32_t* __restrict__ pos_of,
                                                                    const int32_t* __restrict__ copies,
                                                                    int32_t* __restrict__ sorted_copies,
                                                                    int64_t* __restrict__ n_out) {
    constexpr int ND_MAX = 1 << kSortMaxDigit;
    constexpr int DPT = ND_MAX / kSortThreads;                    // digits per thread in the CTA scan (4)
    static_assert(kSortWarps == 8, "the eight per-warp counters of a digit are one 16-byte shared-memory vector");
    __shared__ __align__(16) uint16_t wcnt[ND_MAX][kSortWarps];   // [digit][warp]: per-warp digit counts, then per-warp local bases
    __shared__ int32_t delta[ND_MAX];                             // global position - local position, per digit
    extern __shared__ __align__(16) unsigned char sort_stage[];   // [kSortChunk] keys, then [kSortChunk] uids (kSortStageBytes)
    uint64_t* keys_s = reinterpret_cast<uint64_t*>(sort_stage);
    uint32_t* uid_s = reinterpret_cast<uint32_t*>(keys_s + kSortChunk);
    __shared__ int32_t warp_sums[kSortWarps + 1];
    __shared__ int32_t warp_tot[kSortWarps + 1];                  // the same block scan over the digit totals
    const int nd = 1 << digit_bits;
    const unsigned mask = (unsigned)nd - 1u;
    const int64_t n = FIRST ? n_static : *n_ptr;
    const int64_t base = (int64_t)blockIdx.x * kSortChunk;
    if (base >= n) return;
    const int wib = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < nd; i += kSortThreads) *reinterpret_cast<uint4*>(&wcnt[i][0]) = make_uint4(0u, 0u, 0u, 0u);
    uint64_t key[kSortRounds];
    uint32_t uid[kSortRounds];
    int rank[kSortRounds];             // rank among the warp's earlier elements with the same digit, or -1 (dead)
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {                       // issue every load of the chunk up front
        int64_t idx = base + wib * kSortWarpChunk + r * 32 + lane_id();
        bool live = idx < n;
        key[r] = live ? keys_in[idx] : 0;
        uid[r] = FIRST ? (uint32_t)idx : (live ? uid_in[idx] : 0u);
        if (FIRST && live) live = len[idx] >= k;
        rank[r] = live ? 0 : -1;
    }
    // ranks inside each round first (eight independent matches) ...
    int within[kSortRounds], group[kSortRounds];
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const unsigned d = rank[r] >= 0 ? ((unsigned)(key[r] >> shift) & mask) : (unsigned)nd + lane_id();   // dead lanes match nobody
        const unsigned peers = __match_any_sync(kFull, d);
        within[r] = __popc(peers & lanemask_lt());
        group[r] = __popc(peers);
    }
    __syncthreads();
    // ... then the running per-digit counts of the warp, round after round
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const bool live = rank[r] >= 0;
        const unsigned d = (unsigned)(key[r] >> shift) & mask;
        if (live) rank[r] = (int)wcnt[d][wib] + within[r];
        __syncwarp();
        if (live && within[r] == 0) wcnt[d][wib] = (uint16_t)(rank[r] + group[r]);
        __syncwarp();
    }
    __syncthreads();
    // CTA scan over (digit, warp): local sorted position of every warp's first element of every digit.
    // The eight counters of a digit are one 16-byte vector: one load and one store per digit.
    {
        int32_t cnt[DPT];
        int32_t mine = 0;
        int32_t tot[DPT];                      // global totals of my digits (hist_scanned is scanned per digit only)
        int32_t tmine = 0;
        uint4 vec[DPT];
#pragma unroll
        for (int x = 0; x < DPT; ++x) {
            const int d = threadIdx.x * DPT + x;
            tot[x] = d < nd ? digit_total[d] : 0;
            tmine += tot[x];
            vec[x] = d < nd ? *reinterpret_cast<const uint4*>(&wcnt[d][0]) : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t w01 = vec[x].x, w23 = vec[x].y, w45 = vec[x].z, w67 = vec[x].w;
            cnt[x] = (int32_t)((w01 & 0xffffu) + (w01 >> 16) + (w23 & 0xffffu) + (w23 >> 16) +
                               (w45 & 0xffffu) + (w45 >> 16) + (w67 & 0xffffu) + (w67 >> 16));
            mine += cnt[x];
        }
        int32_t inc = mine, tinc = tmine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t v = __shfl_up_sync(kFull, inc, o);
            int32_t tv = __shfl_up_sync(kFull, tinc, o);
            if ((int)lane_id() >= o) { inc += v; tinc += tv; }
        }
        if (lane_id() == 31) { warp_sums[wib] = inc; warp_tot[wib] = tinc; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int32_t run = 0, trun = 0;
            for (int w = 0; w < kSortWarps; ++w) {
                int32_t t = warp_sums[w]; warp_sums[w] = run; run += t;
                int32_t tt = warp_tot[w]; warp_tot[w] = trun; trun += tt;
            }
            warp_sums[kSortWarps] = run;                           // live elements of the CTA
            warp_tot[kSortWarps] = trun;                           // live elements of the whole input
            if (FIRST && n_out != nullptr && blockIdx.x == 0) *n_out = (int64_t)trun;      // reads with len >= k
            if (table_out != nullptr && blockIdx.x == 0) table_out[nd] = trun;
        }
        __syncthreads();
        int32_t lstart = inc - mine + warp_sums[wib];
        int32_t gbase = tinc - tmine + warp_tot[wib];              // first sorted position of my first digit
#pragma unroll
        for (int x = 0; x < DPT; ++x) {
            const int d = threadIdx.x * DPT + x;
            if (d < nd) {
                // one pass over the whole key: the digit's first position IS the bucket table entry
                if (table_out != nullptr && blockIdx.x == 0) table_out[d] = gbase;
                delta[d] = gbase + hist_scanned[(int64_t)d * C + blockIdx.x] - lstart;
                gbase += tot[x];
                // exclusive prefix over the eight warps, starting at the digit's local start
                uint32_t c[8] = {vec[x].x & 0xffffu, vec[x].x >> 16, vec[x].y & 0xffffu, vec[x].y >> 16,
                                 vec[x].z & 0xffffu, vec[x].z >> 16, vec[x].w & 0xffffu, vec[x].w >> 16};
                uint32_t run = (uint32_t)lstart, b[8];
#pragma unroll
                for (int w = 0; w < 8; ++w) { b[w] = run; run += c[w]; }
                *reinterpret_cast<uint4*>(&wcnt[d][0]) = make_uint4(b[0] | (b[1] << 16), b[2] | (b[3] << 16),
                                                                    b[4] | (b[5] << 16), b[6] | (b[7] << 16));
                lstart += cnt[x];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        if (rank[r] < 0) continue;
        const unsigned d = (unsigned)(key[r] >> shift) & mask;
        const int lp = (int)wcnt[d][wib] + rank[r];
        OVL_CHECK(lp >= 0 && lp < warp_sums[kSortWarps] && lp < kSortChunk);
        keys_s[lp] = key[r];
        uid_s[lp] = uid[r];
    }
    __syncthreads();
    const int live_cta = warp_sums[kSortWarps];
    for (int j = threadIdx.x; j < live_cta; j += kSortThreads) {
        const uint64_t kk = keys_s[j];
        const uint32_t uu = uid_s[j];
        const int pos = delta[(unsigned)(kk >> shift) & mask] + j;
        OVL_CHECK(pos >= 0 && pos < warp_tot[kSortWarps]);
        keys_out[pos] = kk;
        uid_out[pos] = uu;
        if (pos_of != nullptr) pos_of[uu] = pos;
        if (sorted_copies != nullptr) sorted_copies[pos] = copies[uu];
    }
}

// table[t] = first sorted position i with (sorted_key[i] >> shift) >= t, for t in [0, 2^table_bits];
// thread i fills the gap between its predecessor's prefix and its own (thread n: the tail).
__global__ void __launch_bounds__(256) bucket_table_kernel(const uint64_t* __restrict__ sorted_key,
                                                           const int64_t* __restrict__ n_ptr, int shift, int table_bits,
                                                           int32_t* __restrict__ table) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = *n_ptr;
    if (i > n) return;
    const int64_t cur = i < n ? (int64_t)(sorted_key[i] >> shift) : ((int64_t)1 << table_bits);
    const int64_t prev = i > 0 ? (int64_t)(sorted_key[i - 1] >> shift) : -1;
    for (int64_t t = prev + 1; t <= cur; ++t) table[t] = (int32_t)i;
}

// copies of the read at sorted position i (0 past the end of the index): scanned into `cum`.
// sorted_copies (written by the last sort pass) when available, else gathered through sorted_uid.
struct SortedCopies {
    const uint32_t* sorted_uid;
    const int32_t* copies;
    const int64_t* n_ptr;
    __device__ __forceinline__ int64_t operator()(int64_t i) const {
        return i < *n_ptr ? (int64_t)copies[sorted_uid[i]] : 0;
    }
};
// the same from the array the last sort pass wrote (zero past the end of the index: the caller cleared it)
struct SortedCopiesArray {
    const int32_t* sorted_copies;
    __device__ __forceinline__ int64_t operator()(int64_t i) const { return (int64_t)sorted_copies[i]; }
};

// One thread per source read a (overlapGraphs.py:43-52): bucket of suffix_key[a] in the sorted prefix
// keys through the direct-address table; `self_rank` is a's own rank inside that bucket (or -1): the
// reference skips read_b == read_a (:52) and reads are unique, so that is the only skip.
// cnt_out[a] = (candidates, edges) -- edges = copies[a] * sum of copies over the candidates.
__global__ void __launch_bounds__(256) join_count_kernel(const uint64_t* __restrict__ suffix_key,
                                                         const uint64_t* __restrict__ prefix_key,
                                                         const int32_t* __restrict__ len, int k, int64_t U,
                                                         const uint64_t* __restrict__ sorted_key,
                                                         const uint32_t* __restrict__ sorted_uid,
                                                         const int64_t* __restrict__ n_indexed,
                                                         const int32_t* __restrict__ table, int shift,
                                                         const int32_t* __restrict__ pos_of,
                                                         const int32_t* __restrict__ copies, const int64_t* __restrict__ cum,
                                                         int32_t* __restrict__ lo_out, int32_t* __restrict__ self_rank,
                                                         int32_t* __restrict__ cnt_single, I64x2* __restrict__ cnt_pair) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= U) return;
    int64_t cnt = 0, ecnt = 0;
    int32_t lo32 = 0, sr = -1;
    if (len[a] >= k) {
        const uint64_t key = suffix_key[a];
        int64_t lo, hi;
        if (table != nullptr) {
            const uint64_t t = key >> shift;
            lo = table[t];
            hi = table[t + 1];
            OVL_CHECK(lo >= 0 && lo <= hi && hi <= *n_indexed);
            if (shift != 0) {                           // the table narrowed the range: finish inside it
                lo = lower_bound<uint64_t>(sorted_key, lo, hi, key);
                hi = upper_bound<uint64_t>(sorted_key, lo, hi, key);
            }
        } else {
            const int64_t n = *n_indexed;
            lo = lower_bound<uint64_t>(sorted_key, 0, n, key);
            hi = upper_bound<uint64_t>(sorted_key, lo, n, key);
        }
        cnt = hi - lo;
        lo32 = (int32_t)lo;
        bool self;                                      // a sits in its own bucket
        if (pos_of != nullptr) {
            const int64_t mine = pos_of[a];             // a's own sorted position: inside [lo, hi) iff its prefix equals its suffix
            self = mine >= lo && mine < hi;
            if (self) sr = (int32_t)(mine - lo);
        } else {
            self = prefix_key[a] == key;
            if (self) sr = (int32_t)(lower_bound<uint32_t>(sorted_uid, lo, hi, (uint32_t)a) - lo);
        }
        if (self) cnt -= 1;
        if (copies != nullptr) {
            const int64_t ca = copies[a];
            ecnt = ca * (cum[hi] - cum[lo] - (self ? ca : 0));
        }
    }
    lo_out[a] = lo32;
    self_rank[a] = sr;
    if (cnt_pair != nullptr) cnt_pair[a] = I64x2{cnt, ecnt};
    else cnt_single[a] = (int32_t)cnt;          // a bucket holds fewer than 2^31 reads: 4-byte counts, scanned into 8-byte offsets
}

// Layout of the `totals` block join_finalize writes (int64 words), read by the host after one sync.
constexpr int kTotalsPairs = 0;       // candidate pairs, all ranks
constexpr int kTotalsEdges = 1;       // edge rows, all ranks
constexpr int kTotalsBad = 2;         // 2-bit packing met a symbol other than A/C/G/T (count of 64-base groups)
constexpr int kTotalsPBegin = 3;      // this rank's slice of the pair list: [p_begin, p_end)
constexpr int kTotalsPEnd = 4;
constexpr int kTotalsIndexed = 5;     // reads in the index (length >= k)
constexpr int kTotalsBounds = 8;      // kTotalsCuts + 1 pair indices cutting the slice into equal parts ...
constexpr int kTotalsCuts = 64;
constexpr int kTotalsEdgeBounds = kTotalsBounds + kTotalsCuts + 1;     // ... and the edge offset of each
constexpr int kTotalsLen = kTotalsEdgeBounds + kTotalsCuts + 1;

// pair index of cut i of the slice [p_begin, p_begin + P): p_begin + floor(P * i / kTotalsCuts), overflow-free
__device__ __forceinline__ int64_t join_cut(int64_t p_begin, int64_t P, int i) {
    return p_begin + P / kTotalsCuts * i + P % kTotalsCuts * i / kTotalsCuts;
}

// Totals, this rank's slice and the cut points, by ONE small CTA: thread i owns cut i and finds the source read
// whose pair range [pair_off[a], pair_off[a+1]) contains it with a binary search over the scanned counts
// (log2 U dependent loads that hit L2: the scan has just written them).  Round 2's first version ran one thread
// per source read ("the thread whose range contains a cut writes it"): 64 MB of reads and 54 us at 8 M reads for
// 65 numbers.
__global__ void __launch_bounds__(128) join_finalize_kernel(JoinEdgeIndex jx, const int32_t* __restrict__ copies, int64_t U,
                                                            const int32_t* __restrict__ bad, const int64_t* __restrict__ n_indexed,
                                                            int rank, int world, int64_t* __restrict__ totals) {
    const int64_t total = jx.pair_off[U];
    const int64_t etotal = jx.edge_base != nullptr ? jx.edge_base[U] : total;
    const int64_t p_begin = total / world * rank + total % world * rank / world;              // == total * rank / world, no overflow
    const int64_t p_end = total / world * (rank + 1) + total % world * (rank + 1) / world;
    const int64_t P = p_end - p_begin;
    const int i = threadIdx.x;
    if (i == 0) {
        totals[kTotalsPairs] = total;
        totals[kTotalsEdges] = etotal;
        totals[kTotalsBad] = bad != nullptr ? (int64_t)*bad : 0;
        totals[kTotalsPBegin] = p_begin;
        totals[kTotalsPEnd] = p_end;
        totals[kTotalsIndexed] = n_indexed != nullptr ? *n_indexed : 0;
    }
    if (i > kTotalsCuts) return;
    const int64_t p = join_cut(p_begin, P, i);
    if (p >= total) {                                   // a cut at the very end of the whole list
        totals[kTotalsBounds + i] = total;
        totals[kTotalsEdgeBounds + i] = etotal;
        return;
    }
    const int64_t a = upper_bound<int64_t>(jx.pair_off, 0, U + 1, p) - 1;      // pair_off[a] <= p < pair_off[a + 1]
    totals[kTotalsBounds + i] = p;
    totals[kTotalsEdgeBounds + i] = jx.edge_base != nullptr ? join_edge_offset(jx, copies, p, (int32_t)a) : p;
}

// ------------------------------------------------------------------ join fill
// One thread per output pair; a CTA covers a contiguous tile of the output so both stores are
// coalesced.  The owning source read is found by binary search over the scanned counts,
// narrowed first to the tile's own [a_lo, a_hi] range (two searches per CTA).
constexpr int kFillThreads = 256;
constexpr int kFillItems = 8;
constexpr int kFillTile = kFillThreads * kFillItems;

__global__ void __launch_bounds__(kFillThreads) join_fill_kernel(const int64_t* __restrict__ pair_off,  // [nA+1]
                                                                 int64_t nA, int64_t a_begin,
                                                                 const int32_t* __restrict__ lo, const int32_t* __restrict__ self_rank,
                                                                 const uint32_t* __restrict__ sorted_uid,
                                                                 int64_t p_begin, int64_t p_count,
                                                                 int32_t* __restrict__ pair_a, int32_t* __restrict__ pair_b) {
    __shared__ int64_t range[2];
    int64_t tile0 = (int64_t)blockIdx.x * kFillTile;
    if (threadIdx.x == 0) {
        int64_t first = p_begin + tile0;
        int64_t last = p_begin + min(tile0 + kFillTile, p_count) - 1;
        range[0] = upper_bound<int64_t>(pair_off, 0, nA + 1, first) - 1;
        range[1] = upper_bound<int64_t>(pair_off, 0, nA + 1, last) - 1;
    }
    __syncthreads();
    int64_t alo = range[0], ahi = range[1];
#pragma unroll
    for (int it = 0; it < kFillItems; ++it) {
        int64_t q = tile0 + it * kFillThreads + threadIdx.x;
        if (q >= p_count) break;
        int64_t p = p_begin + q;
        int64_t i = upper_bound<int64_t>(pair_off, alo, ahi + 1, p) - 1;
        int32_t r = (int32_t)(p - pair_off[i]);
        int32_t sr = self_rank[i];
        if (sr >= 0 && r >= sr) r += 1;
        pair_a[q] = (int32_t)(a_begin + i);
        pair_b[q] = (int32_t)sorted_uid[lo[i] + r];
    }
}

// Same output, LANES lanes per source read (32: a warp streams a large bucket; 8 / 4 / 1 for smaller
// buckets, so that lanes are not idle): no search, the lanes of a group write consecutive pairs; four
// independent loads in flight per lane.
template <int LANES>
__global__ void __launch_bounds__(256) join_fill_group_kernel(const int64_t* __restrict__ pair_off, int64_t nA, int64_t a_begin,
                                                              const int32_t* __restrict__ lo, const int32_t* __restrict__ self_rank,
                                                              const uint32_t* __restrict__ sorted_uid,
                                                              int64_t p_begin, int64_t p_count,
                                                              int32_t* __restrict__ pair_a, int32_t* __restrict__ pair_b) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / LANES;
    if (i >= nA) return;
    const int sub = (int)(t % LANES);
    const int64_t first = pair_off[i], last = pair_off[i + 1];
    const int64_t from = max(first, p_begin), to = min(last, p_begin + p_count);
    if (from >= to) return;
    const int32_t sr = self_rank[i];
    const uint32_t* __restrict__ bucket = sorted_uid + lo[i];
    const int32_t a = (int32_t)(a_begin + i);
    int64_t p = from + sub;
    for (; p + 3 * LANES < to; p += 4 * LANES) {
        int32_t b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int32_t r = (int32_t)(p + LANES * j - first);
            if (sr >= 0 && r >= sr) r += 1;
            OVL_CHECK(r >= 0 && (int64_t)r <= last - first);        // the bucket holds the candidates and, at most, the read itself
            b[j] = (int32_t)bucket[r];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t q = p + LANES * j - p_begin;
            OVL_CHECK(q >= 0 && q < p_count);
            __stcs(pair_a + q, a);                      // streaming: the GBs of pairs must not evict the index from L2
            __stcs(pair_b + q, b[j]);
        }
    }
    for (; p < to; p += LANES) {
        int32_t r = (int32_t)(p - first);
        if (sr >= 0 && r >= sr) r += 1;
        int64_t q = p - p_begin;
        OVL_CHECK(q >= 0 && q < p_count);
        __stcs(pair_a + q, a);
        __stcs(pair_b + q, (int32_t)bucket[r]);
    }
}

// k == 0: every ordered pair a != b (overlapGraphs.py:49), a in [a_begin, a_end).
__global__ void __launch_bounds__(256) all_pairs_fill_kernel(int64_t U, int64_t a_begin, int64_t p_begin, int64_t p_count,
                                                             int32_t* __restrict__ pair_a, int32_t* __restrict__ pair_b) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= p_count) return;
    int64_t p = p_begin + q;
    int64_t per = U - 1;
    int64_t ai = p / per;
    int64_t r = p - ai * per;
    int64_t a = a_begin + ai;
    pair_a[q] = (int32_t)a;
    pair_b[q] = (int32_t)(r >= a ? r + 1 : r);
}

}  // namespace ovl
