// k-mer stages of the overlap graph builder (replaces overlapGraphs.py:30-52, 55-60):
//   K0 pack_reads      ASCII -> 2-bit packed rows (+ length, non-ACGT detection)
//   K1 kmer_keys       prefix / suffix k-mer of every read as a 2k-bit integer
//   (K2 index and K3 join live in index.cuh)
//   K6 expand_*        (a, b, score, end) -> copy_a x copy_b edge rows
// All of these are HBM-bound byte/integer work: coalesced, 128-bit where the layout allows.
#pragma once
#include "common.cuh"

namespace ovl {

// Base code: (c >> 1) & 3  ->  A=0, C=1, T=2, G=3.  Any bijection works: keys are only
// compared for equality and the DP only tests s[i] == t[j].
constexpr uint64_t kInvalidKey = ~0ull;

// ------------------------------------------------------------------ K0 pack_reads
// One thread per 4 output words (64 bases, one 16-byte store).  The read's ASCII bytes start at
// an arbitrary byte offset, so the thread loads the five aligned 16-byte segments that cover
// its 64 bytes (all issued up front) and funnel-shifts them into place; the segment shared with
// the neighbouring thread comes from L1, DRAM sees each byte once.  Requires: ascii base 16-byte
// aligned and >= 32 bytes of slack after the last read.
__device__ __forceinline__ uint32_t pack16(const uint32_t r[4], int nvalid, uint32_t& bad) {
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int nv = nvalid - 4 * i;                     // valid bytes in this word
        uint32_t keep = nv >= 4 ? 0xffffffffu : (nv <= 0 ? 0u : ((1u << (8 * nv)) - 1u));
        uint32_t c = (r[i] & keep) | (0x41414141u & ~keep);          // filler 'A' -> code 0
        uint32_t t = (c >> 1) & 0x03030303u;
        // exact check: rebuild the byte each code stands for and compare
        uint32_t is_t = (t >> 1) & ~t & 0x01010101u;                 // code 2 == 'T'
        uint32_t expect = 0x41414141u + 2u * t + 0x0fu * is_t;       // A 41, C 43, G 47, T 54
        bad |= c ^ expect;
        // gather the four 2-bit fields (at bits 0,8,16,24) into one byte with a multiply
        uint32_t pk = (t * 0x01041040u) >> 24;
        out |= pk << (8 * i);
    }
    return out;
}

// bits [o, o + 64) of the word array w[0..7] (o < 192), without dynamic register indexing
__device__ __forceinline__ uint64_t window64(const uint32_t (&w)[8], int o) {
    const int wi = o >> 5, sh = o & 31;
    uint32_t a = 0, b = 0, c = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j)
        if (wi == j) { a = w[j]; b = w[j + 1]; c = w[j + 2]; }
    return ((uint64_t)__funnelshift_r(b, c, sh) << 32) | __funnelshift_r(a, b, sh);
}

// KEYS: K1 fused into K0 -- the thread that packs a read's first 64 bases holds its prefix k-mer, the
// thread that packs its last bases holds (with one shuffle from its left neighbour when the k-mer
// straddles two 64-base groups) its suffix k-mer, so the keys cost no further pass over the rows.
template <bool KEYS>
__global__ void __launch_bounds__(256) pack_reads_kernel(const uint8_t* __restrict__ ascii,
                                                         const int64_t* __restrict__ offsets,
                                                         int64_t U, int row_words,
                                                         uint32_t* __restrict__ packed,
                                                         int32_t* __restrict__ len_out,
                                                         int32_t* __restrict__ bad_count,
                                                         int k, const int32_t* __restrict__ segment,
                                                         uint64_t* __restrict__ prefix_key,
                                                         uint64_t* __restrict__ suffix_key) {
    const int quads = row_words >> 2;                // 16-byte groups per row
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = slot < U * quads;
    if (!KEYS && !active) return;
    const int64_t u = active ? slot / quads : 0;
    const int q = active ? (int)(slot - u * quads) : 0;
    int64_t o0 = 0;
    int len = 0;
    if (active) {
        o0 = offsets[u];
        len = (int)(offsets[u + 1] - o0);
        if (q == 0) len_out[u] = len;
    }
    const int nvalid = len - 64 * q;                 // bases this thread holds
    uint4 outv = make_uint4(0u, 0u, 0u, 0u);
    if (nvalid > 0) {
        int64_t addr = o0 + 64 * (int64_t)q;         // byte index of the first base
        const uint4* seg = reinterpret_cast<const uint4*>(ascii + (addr & ~(int64_t)15));
        int nseg = min(5, (int)(((addr & 15) + min(nvalid, 64) + 15) >> 4));
        uint4 sv[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) sv[i] = i < nseg ? __ldg(seg + i) : make_uint4(0u, 0u, 0u, 0u);
        uint32_t x[20];
#pragma unroll
        for (int i = 0; i < 5; ++i) { x[4 * i] = sv[i].x; x[4 * i + 1] = sv[i].y; x[4 * i + 2] = sv[i].z; x[4 * i + 3] = sv[i].w; }
        unsigned sh = (unsigned)(addr & 15);
        // shift right by sh bytes: 8, 4, then 0..3 bytes
        uint32_t y[18], z[17], r[16];
#pragma unroll
        for (int i = 0; i < 18; ++i) y[i] = (sh & 8) ? x[i + 2] : x[i];
#pragma unroll
        for (int i = 0; i < 17; ++i) z[i] = (sh & 4) ? y[i + 1] : y[i];
        unsigned bs = (sh & 3) * 8;
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __funnelshift_r(z[i], z[i + 1], bs);
        uint32_t bad = 0;
        outv.x = pack16(r, nvalid, bad);
        outv.y = pack16(r + 4, nvalid - 16, bad);
        outv.z = pack16(r + 8, nvalid - 32, bad);
        outv.w = pack16(r + 12, nvalid - 48, bad);
        if (bad) atomicAdd(bad_count, 1);
    }
    if (active) reinterpret_cast<uint4*>(packed)[slot] = outv;
    if (KEYS) {
        // bases 32..63 of the left neighbour's group (the same read's previous group when q > 0)
        const uint32_t pz = __shfl_up_sync(kFull, outv.z, 1), pw = __shfl_up_sync(kFull, outv.w, 1);
        if (!active) return;
        const uint64_t mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
        const uint64_t tag = segment != nullptr ? (uint64_t)(uint32_t)segment[u] << (2 * k) : 0ull;
        if (q == 0)                                                       // overlapGraphs.py:33-37: prefix = read[:k]
            prefix_key[u] = len >= k ? (((((uint64_t)outv.y << 32) | outv.x) & mask) | tag) : kInvalidKey;
        if (q == (len > 0 ? (len - 1) >> 6 : 0)) {                        // :44-47: suffix = read[-k:]
            uint64_t sk = kInvalidKey;
            if (len >= k) {
                const int s_local = len - k - 64 * q;                     // first base of the k-mer, relative to my group (>= -31)
                if (s_local >= 0 || lane_id() != 0) {
                    const uint32_t w[8] = {pz, pw, outv.x, outv.y, outv.z, outv.w, 0u, 0u};
                    sk = window64(w, 64 + 2 * s_local);
                } else {
                    // the k-mer starts in a group packed by another warp: take it from the ASCII bytes (L1/L2 hits)
                    sk = 0;
                    const uint8_t* p = ascii + o0 + (len - k);
                    for (int i = 0; i < k; ++i) sk |= (uint64_t)((p[i] >> 1) & 3u) << (2 * i);
                }
                sk = (sk & mask) | tag;
            }
            suffix_key[u] = sk;
        }
    }
}

// ------------------------------------------------------------------ K1 kmer_keys
// overlapGraphs.py:33-37 (prefix = read[:k]) and :44-47 (suffix = read[-k:]); reads shorter
// than k can only ever match themselves: they get a placeholder key and are kept
// out of the index and the lookups by their length (a 32-mer of all G is a real key ~0).
__device__ __forceinline__ uint64_t extract_bits64(const uint32_t* __restrict__ row, int row_words, int bitpos) {
    int w = bitpos >> 5, s = bitpos & 31;
    uint32_t a = row[w];
    uint32_t b = w + 1 < row_words ? row[w + 1] : 0u;
    uint32_t c = w + 2 < row_words ? row[w + 2] : 0u;
    uint32_t lo = __funnelshift_r(a, b, s);
    uint32_t hi = __funnelshift_r(b, c, s);
    return ((uint64_t)hi << 32) | lo;
}

// `segment` (optional) tags every read with the read set it belongs to; the tag goes into the key
// bits above the k-mer, so that one index / one join serves many independent read sets at once
// (the parameter sweep of experiments.py) without ever pairing reads across sets.
__global__ void __launch_bounds__(256) kmer_keys_kernel(const uint32_t* __restrict__ packed, int row_words,
                                                        const int32_t* __restrict__ len, int64_t U, int k,
                                                        const int32_t* __restrict__ segment,
                                                        uint64_t* __restrict__ prefix_key,
                                                        uint64_t* __restrict__ suffix_key) {
    int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    int n = len[u];
    uint64_t pk = kInvalidKey, sk = kInvalidKey;
    if (n >= k) {
        const uint32_t* row = packed + u * row_words;
        uint64_t mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
        pk = (((uint64_t)row[1] << 32) | row[0]) & mask;
        sk = extract_bits64(row, row_words, 2 * (n - k)) & mask;
        if (segment != nullptr) {
            uint64_t tag = (uint64_t)(uint32_t)segment[u] << (2 * k);
            pk |= tag;
            sk |= tag;
        }
    }
    prefix_key[u] = pk;
    suffix_key[u] = sk;
}

// ------------------------------------------------------------------ K1h / K3v: k > 32
// A k-mer longer than 32 bases does not fit a u64, so the index is built on a 64-bit HASH of the
// k-mer and the join verifies every candidate by comparing the two k-mers base by base (2-bit
// words): equal hashes are necessary, the comparison makes the result exact.  Buckets are tiny at
// such k, so one thread per source read walks its hash bucket.
__device__ __forceinline__ uint64_t kmer_word(const uint32_t* __restrict__ row, int row_words, int start, int k, int w) {
    // bases [start + 32w, start + 32w + 32) of the read, clipped to the k-mer, as one 64-bit word
    uint64_t v = extract_bits64(row, row_words, 2 * (start + 32 * w));
    int left = k - 32 * w;
    return left >= 32 ? v : (v & ((1ull << (2 * left)) - 1ull));
}
__device__ __forceinline__ uint64_t kmer_hash(const uint32_t* __restrict__ row, int row_words, int start, int k) {
    uint64_t h = 0x243f6a8885a308d3ull ^ (uint64_t)k;
    for (int w = 0; 32 * w < k; ++w) {
        h ^= kmer_word(row, row_words, start, k, w);
        h *= 0x9e3779b97f4a7c15ull;
        h ^= h >> 32;
    }
    return h;
}
__device__ __forceinline__ bool kmer_equal(const uint32_t* __restrict__ ra, int sa, const uint32_t* __restrict__ rb, int sb,
                                           int row_words, int k) {
    for (int w = 0; 32 * w < k; ++w)
        if (kmer_word(ra, row_words, sa, k, w) != kmer_word(rb, row_words, sb, k, w)) return false;
    return true;
}

__global__ void __launch_bounds__(256) kmer_hash_kernel(const uint32_t* __restrict__ packed, int row_words,
                                                        const int32_t* __restrict__ len, int64_t U, int k,
                                                        uint64_t* __restrict__ prefix_key, uint64_t* __restrict__ suffix_key) {
    int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    int n = len[u];
    uint64_t pk = 0, sk = 0;
    if (n >= k) {
        const uint32_t* row = packed + u * row_words;
        pk = kmer_hash(row, row_words, 0, k);
        sk = kmer_hash(row, row_words, n - k, k);
    }
    prefix_key[u] = pk;
    suffix_key[u] = sk;
}

// count (FILL = false) or write (FILL = true) the verified candidates of every source read
template <bool FILL>
__global__ void __launch_bounds__(256) join_verify_kernel(const uint32_t* __restrict__ packed, int row_words,
                                                          const int32_t* __restrict__ len, int k,
                                                          const uint64_t* __restrict__ suffix_key, int64_t nA, int64_t a_begin,
                                                          const uint64_t* __restrict__ sorted_key, const uint32_t* __restrict__ sorted_uid,
                                                          const int64_t* __restrict__ n_indexed,
                                                          int64_t* __restrict__ cnt_out,            // !FILL
                                                          const int64_t* __restrict__ pair_off,      // FILL
                                                          int64_t p_begin, int64_t p_count,
                                                          int32_t* __restrict__ pair_a, int32_t* __restrict__ pair_b) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nA) return;
    int64_t a = a_begin + i;
    int64_t cnt = 0;
    int na = len[a];
    if (na >= k) {
        uint64_t key = suffix_key[a];
        int64_t n = *n_indexed;
        int64_t lo = lower_bound<uint64_t>(sorted_key, 0, n, key);
        const uint32_t* ra = packed + a * row_words;
        int64_t out = FILL ? pair_off[i] : 0;
        for (int64_t j = lo; j < n && sorted_key[j] == key; ++j) {
            int64_t b = sorted_uid[j];
            if (b == a) continue;                                        // overlapGraphs.py:52
            if (!kmer_equal(ra, na - k, packed + b * row_words, 0, row_words, k)) continue;   // hash collision
            if (FILL) {
                int64_t q = out + cnt - p_begin;
                if (q >= 0 && q < p_count) { pair_a[q] = (int32_t)a; pair_b[q] = (int32_t)b; }
            }
            ++cnt;
        }
    }
    if (!FILL) cnt_out[i] = cnt;
}

// ------------------------------------------------------------------ byte-coded reads (any alphabet)
// Read sets with more than four distinct symbols cannot be 2-bit packed.  They are kept as padded
// byte rows (row_bytes per read, 16-byte aligned) and take the general route: hashed k-mer keys with
// byte-wise verification in the join, and the CTA-per-pair anti-diagonal DP (dp.cuh) comparing bytes.
__global__ void __launch_bounds__(256) pack_bytes_kernel(const uint8_t* __restrict__ ascii, const int64_t* __restrict__ offsets,
                                                         int64_t U, int row_bytes, uint8_t* __restrict__ rows,
                                                         int32_t* __restrict__ len_out) {
    int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per 4 output bytes
    int quads = row_bytes >> 2;
    if (slot >= U * quads) return;
    int64_t u = slot / quads;
    int q = (int)(slot - u * quads);
    int64_t o0 = offsets[u];
    int len = (int)(offsets[u + 1] - o0);
    if (q == 0) len_out[u] = len;
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        int i = 4 * q + b;
        if (i < len) w |= (uint32_t)ascii[o0 + i] << (8 * b);
    }
    reinterpret_cast<uint32_t*>(rows)[slot] = w;
}

__device__ __forceinline__ uint64_t bytes_hash(const uint8_t* __restrict__ p, int k) {
    uint64_t h = 0x243f6a8885a308d3ull ^ (uint64_t)k;
    for (int i = 0; i < k; ++i) { h ^= p[i]; h *= 0x100000001b3ull; h ^= h >> 29; }
    return h * 0x9e3779b97f4a7c15ull;
}

__global__ void __launch_bounds__(256) kmer_hash8_kernel(const uint8_t* __restrict__ rows, int row_bytes,
                                                         const int32_t* __restrict__ len, int64_t U, int k,
                                                         uint64_t* __restrict__ prefix_key, uint64_t* __restrict__ suffix_key) {
    int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    int n = len[u];
    uint64_t pk = 0, sk = 0;
    if (n >= k) {
        const uint8_t* row = rows + u * row_bytes;
        pk = bytes_hash(row, k);
        sk = bytes_hash(row + (n - k), k);
    }
    prefix_key[u] = pk;
    suffix_key[u] = sk;
}

template <bool FILL>
__global__ void __launch_bounds__(256) join_verify8_kernel(const uint8_t* __restrict__ rows, int row_bytes,
                                                           const int32_t* __restrict__ len, int k,
                                                           const uint64_t* __restrict__ suffix_key, int64_t nA, int64_t a_begin,
                                                           const uint64_t* __restrict__ sorted_key, const uint32_t* __restrict__ sorted_uid,
                                                           const int64_t* __restrict__ n_indexed, int64_t* __restrict__ cnt_out,
                                                           const int64_t* __restrict__ pair_off, int64_t p_begin, int64_t p_count,
                                                           int32_t* __restrict__ pair_a, int32_t* __restrict__ pair_b) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nA) return;
    int64_t a = a_begin + i;
    int64_t cnt = 0;
    int na = len[a];
    if (na >= k) {
        uint64_t key = suffix_key[a];
        int64_t n = *n_indexed;
        int64_t lo = lower_bound<uint64_t>(sorted_key, 0, n, key);
        const uint8_t* sa = rows + a * row_bytes + (na - k);
        int64_t out = FILL ? pair_off[i] : 0;
        for (int64_t j = lo; j < n && sorted_key[j] == key; ++j) {
            int64_t b = sorted_uid[j];
            if (b == a) continue;
            const uint8_t* pb = rows + b * row_bytes;
            bool same = true;
            for (int x = 0; x < k; ++x) if (sa[x] != pb[x]) { same = false; break; }
            if (!same) continue;
            if (FILL) {
                int64_t q = out + cnt - p_begin;
                if (q >= 0 && q < p_count) { pair_a[q] = (int32_t)a; pair_b[q] = (int32_t)b; }
            }
            ++cnt;
        }
    }
    if (!FILL) cnt_out[i] = cnt;
}

// ------------------------------------------------------------------ K6 expand edges
// Edge row = (node_a, node_b, weight, end_position), node id = node_off[uid] + copy, emitted
// in the reference's insertion order: pair order, then copy_a, then copy_b (overlapGraphs.py:55-60).
__global__ void __launch_bounds__(kFillThreads) expand_fill_kernel(const int64_t* __restrict__ edge_off,  // [P+1]
                                                                   int64_t P,
                                                                   const int32_t* __restrict__ pair_a, const int32_t* __restrict__ pair_b,
                                                                   const int32_t* __restrict__ score, const int32_t* __restrict__ end,
                                                                   const int32_t* __restrict__ copies, const int64_t* __restrict__ node_off,
                                                                   int64_t e_begin, int64_t e_count, int4* __restrict__ edges) {
    __shared__ int64_t range[2];
    int64_t tile0 = (int64_t)blockIdx.x * kFillTile;
    if (threadIdx.x == 0) {
        int64_t first = e_begin + tile0;
        int64_t last = e_begin + min(tile0 + kFillTile, e_count) - 1;
        range[0] = upper_bound<int64_t>(edge_off, 0, P + 1, first) - 1;
        range[1] = upper_bound<int64_t>(edge_off, 0, P + 1, last) - 1;
    }
    __syncthreads();
    int64_t plo = range[0], phi = range[1];
#pragma unroll
    for (int it = 0; it < kFillItems; ++it) {
        int64_t q = tile0 + it * kFillThreads + threadIdx.x;
        if (q >= e_count) break;
        int64_t e = e_begin + q;
        int64_t p = upper_bound<int64_t>(edge_off, plo, phi + 1, e) - 1;
        int64_t r = e - edge_off[p];
        int32_t a = pair_a[p], b = pair_b[p];
        int32_t cb = copies[b];
        int32_t ia = (int32_t)(r / cb), ib = (int32_t)(r - (int64_t)ia * cb);
        edges[q] = make_int4((int32_t)(node_off[a] + ia), (int32_t)(node_off[b] + ib), score[p], end[p]);
    }
}

// order-preserving compaction of the edge rows with weight >= min_weight (keep_off = exclusive scan
// of the keep flags): the `if score > 0` of the all-pairs builders, overlapGraphs.py:225 and :347
__global__ void __launch_bounds__(256) filter_edges_kernel(const int4* __restrict__ edges, const int64_t* __restrict__ keep_off,
                                                           int64_t E, int32_t min_weight, int4* __restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int4 row = edges[e];
    if (row.z >= min_weight) out[keep_off[e]] = row;
}

// every read appears once: edge row == pair row, no scan needed
__global__ void __launch_bounds__(256) expand_unit_kernel(const int32_t* __restrict__ pair_a, const int32_t* __restrict__ pair_b,
                                                          const int32_t* __restrict__ score, const int32_t* __restrict__ end,
                                                          int64_t P, int4* __restrict__ edges) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    edges[p] = make_int4(pair_a[p], pair_b[p], score[p], end[p]);
}

}  // namespace ovl
