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

constexpr int kSortWarps = 4;            // warps per CTA
constexpr int kSortThreads = kSortWarps * 32;
constexpr int kSortChunk = 2048;         // consecutive elements owned by one warp
constexpr int kSortBatch = 16;           // loads in flight per lane
constexpr int kSortMaxDigit = 10;        // bits per pass: 4 warps x 1024 counters = 16 KB of shared memory
constexpr int kTableMaxBits = 22;        // direct-address table: at most 4 Mi + 1 entries (16 MB)

__host__ __device__ inline int sort_passes(int key_bits) { return (key_bits + kSortMaxDigit - 1) / kSortMaxDigit; }
__host__ __device__ inline int sort_digit_bits(int key_bits) {
    int p = sort_passes(key_bits);
    return (key_bits + p - 1) / p;
}

// per-warp digit histogram of the warp's chunk -> hist[digit * W + warp]
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const uint64_t* __restrict__ keys,
                                                                 const int32_t* __restrict__ len, int k,
                                                                 const int64_t* __restrict__ n_ptr, int64_t n_static,
                                                                 int shift, int digit_bits, int64_t W,
                                                                 int32_t* __restrict__ hist) {
    __shared__ int32_t cnt[kSortWarps][1 << kSortMaxDigit];
    const int wib = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * kSortWarps + wib;
    const int nd = 1 << digit_bits;
    const unsigned mask = (unsigned)nd - 1u;
    for (int i = lane_id(); i < nd; i += 32) cnt[wib][i] = 0;
    __syncwarp();
    if (warp >= W) return;
    const int64_t n = FIRST ? n_static : *n_ptr;
    const int64_t base = warp * kSortChunk;
    for (int b = 0; b < kSortChunk / 32; b += kSortBatch) {
        if (base + (int64_t)b * 32 >= n) break;
        uint64_t key[kSortBatch];
        bool live[kSortBatch];
#pragma unroll
        for (int it = 0; it < kSortBatch; ++it) {                 // all loads of the batch before the first atomic
            int64_t idx = base + (int64_t)(b + it) * 32 + lane_id();
            live[it] = idx < n;
            key[it] = live[it] ? keys[idx] : 0;
            if (FIRST && live[it]) live[it] = len[idx] >= k;
        }
#pragma unroll
        for (int it = 0; it < kSortBatch; ++it)
            if (live[it]) atomicAdd(&cnt[wib][(unsigned)(key[it] >> shift) & mask], 1);
    }
    __syncwarp();
    for (int d = lane_id(); d < nd; d += 32) hist[(int64_t)d * W + warp] = cnt[wib][d];
}

// stable scatter of the warp's chunk to the scanned offsets.  FIRST: input is (prefix_key, uid = index)
// straight from K1 and reads shorter than k are dropped; n_out receives the number of survivors.
// pos_of (last pass only): sorted position of every uid.
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const uint64_t* __restrict__ keys_in,
                                                                    const uint32_t* __restrict__ uid_in,
                                                                    const int32_t* __restrict__ len, int k,
                                                                    const int64_t* __restrict__ n_ptr, int64_t n_static,
                                                                    int shift, int digit_bits, int64_t W,
                                                                    const int32_t* __restrict__ hist_scanned,
                                                                    uint64_t* __restrict__ keys_out,
                                                                    uint32_t* __restrict__ uid_out,
                                                                    int32_t* __restrict__ pos_of,
                                                                    int64_t* __restrict__ n_out) {
    __shared__ int32_t off[kSortWarps][1 << kSortMaxDigit];
    const int wib = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * kSortWarps + wib;
    if (warp >= W) return;
    const int nd = 1 << digit_bits;
    const unsigned mask = (unsigned)nd - 1u;
    for (int d = lane_id(); d < nd; d += 32) off[wib][d] = hist_scanned[(int64_t)d * W + warp];
    __syncwarp();
    const int64_t n = FIRST ? n_static : *n_ptr;
    const int64_t base = warp * kSortChunk;
    for (int b = 0; b < kSortChunk / 32; b += kSortBatch) {
        if (base + (int64_t)b * 32 >= n) break;
        uint64_t key[kSortBatch];
        uint32_t uid[kSortBatch];
        bool live[kSortBatch];
#pragma unroll
        for (int it = 0; it < kSortBatch; ++it) {                 // issue every load of the batch up front
            int64_t idx = base + (int64_t)(b + it) * 32 + lane_id();
            live[it] = idx < n;
            key[it] = live[it] ? keys_in[idx] : 0;
            uid[it] = FIRST ? (uint32_t)idx : (live[it] ? uid_in[idx] : 0u);
            if (FIRST && live[it]) live[it] = len[idx] >= k;
        }
#pragma unroll
        for (int it = 0; it < kSortBatch; ++it) {
            unsigned d = live[it] ? ((unsigned)(key[it] >> shift) & mask) : (unsigned)nd + lane_id();   // dead lanes match nobody
            unsigned peers = __match_any_sync(kFull, d);
            int rank = __popc(peers & lanemask_lt());
            int pos = 0;
            if (live[it]) pos = off[wib][d] + rank;
            __syncwarp();
            if (live[it] && rank == 0) off[wib][d] += __popc(peers);
            __syncwarp();
            if (live[it]) {
                keys_out[pos] = key[it];
                uid_out[pos] = uid[it];
                if (pos_of != nullptr) pos_of[uid[it]] = pos;
            }
        }
    }
    if (FIRST && n_out != nullptr && warp == W - 1 && lane_id() == 0) {
        // after the last warp's chunk, the running offset of the last digit is the number of survivors
        *n_out = (int64_t)off[wib][nd - 1];
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

// copies of the read at sorted position i (0 past the end of the index): scanned into `cum`
struct SortedCopies {
    const uint32_t* sorted_uid;
    const int32_t* copies;
    const int64_t* n_ptr;
    __device__ __forceinline__ int64_t operator()(int64_t i) const { return i < *n_ptr ? (int64_t)copies[sorted_uid[i]] : 0; }
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
                                                         int64_t* __restrict__ cnt_single, I64x2* __restrict__ cnt_pair) {
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
        bool self = prefix_key[a] == key;               // a sits in its own bucket
        if (self) {
            sr = pos_of != nullptr ? pos_of[a] - lo32 : (int32_t)(lower_bound<uint32_t>(sorted_uid, lo, hi, (uint32_t)a) - lo);
            cnt -= 1;
        }
        if (copies != nullptr) {
            const int64_t ca = copies[a];
            ecnt = ca * (cum[hi] - cum[lo] - (self ? ca : 0));
        }
    }
    lo_out[a] = lo32;
    self_rank[a] = sr;
    if (cnt_pair != nullptr) cnt_pair[a] = I64x2{cnt, ecnt};
    else cnt_single[a] = cnt;
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

__device__ __forceinline__ int64_t join_edge_offset_at(const JoinEdgeIndex& jx, const int32_t* __restrict__ copies, int64_t U,
                                                       int64_t p, int64_t total_pairs, int64_t total_edges) {
    if (jx.edge_base == nullptr) return p;              // every read occurs once: one row per pair
    if (p >= total_pairs) return total_edges;
    int64_t a = upper_bound<int64_t>(jx.pair_off, 0, U + 1, p) - 1;      // the (non-empty) source that owns pair p
    return join_edge_offset(jx, copies, p, (int32_t)a);
}

__global__ void __launch_bounds__(128) join_finalize_kernel(JoinEdgeIndex jx, const int32_t* __restrict__ copies, int64_t U,
                                                            const int32_t* __restrict__ bad, const int64_t* __restrict__ n_indexed,
                                                            int rank, int world, int64_t* __restrict__ totals) {
    const int64_t total = jx.pair_off[U];
    const int64_t etotal = jx.edge_base != nullptr ? jx.edge_base[U] : total;
    const int64_t p_begin = total / world * rank + total % world * rank / world;       // == total * rank / world, no overflow
    const int64_t p_end = total / world * (rank + 1) + total % world * (rank + 1) / world;
    const int i = threadIdx.x;
    if (i == 0) {
        totals[kTotalsPairs] = total;
        totals[kTotalsEdges] = etotal;
        totals[kTotalsBad] = bad != nullptr ? (int64_t)*bad : 0;
        totals[kTotalsPBegin] = p_begin;
        totals[kTotalsPEnd] = p_end;
        totals[kTotalsIndexed] = n_indexed != nullptr ? *n_indexed : 0;
    }
    if (i <= kTotalsCuts) {
        const int64_t P = p_end - p_begin;
        const int64_t p = p_begin + P / kTotalsCuts * i + P % kTotalsCuts * i / kTotalsCuts;
        totals[kTotalsBounds + i] = p;
        totals[kTotalsEdgeBounds + i] = join_edge_offset_at(jx, copies, U, p, total, etotal);
    }
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

// Same output, one warp per source read: used when buckets are large (mean >= 32 candidates per
// read), where every lane streams consecutive candidates -- no search, fully coalesced stores; four
// independent loads in flight per lane.
__global__ void __launch_bounds__(256) join_fill_warp_kernel(const int64_t* __restrict__ pair_off, int64_t nA, int64_t a_begin,
                                                             const int32_t* __restrict__ lo, const int32_t* __restrict__ self_rank,
                                                             const uint32_t* __restrict__ sorted_uid,
                                                             int64_t p_begin, int64_t p_count,
                                                             int32_t* __restrict__ pair_a, int32_t* __restrict__ pair_b) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nA) return;
    int64_t first = pair_off[i], last = pair_off[i + 1];
    int64_t from = max(first, p_begin), to = min(last, p_begin + p_count);
    if (from >= to) return;
    const int32_t sr = self_rank[i];
    const uint32_t* __restrict__ bucket = sorted_uid + lo[i];
    const int32_t a = (int32_t)(a_begin + i);
    int64_t p = from + lane_id();
    for (; p + 96 < to; p += 128) {
        int32_t b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int32_t r = (int32_t)(p + 32 * j - first);
            if (sr >= 0 && r >= sr) r += 1;
            b[j] = (int32_t)bucket[r];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t q = p + 32 * j - p_begin;
            pair_a[q] = a;
            pair_b[q] = b[j];
        }
    }
    for (; p < to; p += 32) {
        int32_t r = (int32_t)(p - first);
        if (sr >= 0 && r >= sr) r += 1;
        int64_t q = p - p_begin;
        pair_a[q] = a;
        pair_b[q] = (int32_t)bucket[r];
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
