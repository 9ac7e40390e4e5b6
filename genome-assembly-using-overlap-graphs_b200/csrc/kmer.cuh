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

// ------------------------------------------------------------------ K0 pack_reads (+ K1 fused)
// One thread per 4 output words (64 bases, one 16-byte store), 256 such slots per CTA.  The slots of a CTA
// cover one contiguous span of the ASCII input (<= 16 KB), so the CTA
//   1. converts the span in ALIGNED 16-byte pieces -- perfectly coalesced 128-bit loads, each piece
//      becomes one 32-bit word of 2-bit codes in shared memory (with an exact A/C/G/T check), and
//   2. lets every thread cut its 4 output words out of that bit stream at its read's own (arbitrary) byte
//      offset: two shared-memory words and one funnel shift per output word.
// The byte re-alignment therefore happens on the 4x smaller packed stream, not on the ASCII bytes.
// KEYS (K1 fused): the prefix k-mer is the first words of the row; the suffix k-mer is cut out of the same
// bit stream (the span starts 32 bases early so that a k-mer reaching back into the previous slot is there).
// Requires: ascii base 16-byte aligned and >= 32 bytes of slack after the last read.
constexpr int kPackThreads = 256;
constexpr int kPackPieces = (kPackThreads * 64 + 32 + 16) / 16 + 2;      // 16-byte pieces a CTA's span can touch
constexpr int kPackIters = (kPackPieces + 3 + kPackThreads - 1) / kPackThreads;      // rounds of the conversion loop (5)
#ifndef OVL_PACK_PRELOADS
#define OVL_PACK_PRELOADS 4          // rounds of the conversion loop whose 16-byte load is issued up front (4: 32 registers, no spill, 435 us at 8 M reads; 5: two spilled words, 454 us; 3: 444 us)
#endif
#ifndef OVL_PACK_MINB
#define OVL_PACK_MINB 8                  // resident CTAs per SM: measured at 8 M reads 455 us (8, two spilled words) / 465 (6) / 496 (5) / 555 (4); the loop without the preloads: 475
#endif

// 4 ASCII bytes -> their 2-bit codes gathered in byte 3 of the result; `bad` collects c ^ (the letter each code stands for)
__device__ __forceinline__ uint32_t codes4(uint32_t c, uint32_t& bad) {
    uint32_t t = (c >> 1) & 0x03030303u;
    uint32_t is_t = (t >> 1) & ~t & 0x01010101u;                 // code 2 == 'T'
    uint32_t expect = 0x41414141u + 2u * t + 0x0fu * is_t;       // A 41, C 43, G 47, T 54
    bad |= c ^ expect;
    return t * 0x01041040u;                                      // the four 2-bit fields land in bits 24..31
}

template <bool KEYS>
__global__ void __launch_bounds__(kPackThreads, OVL_PACK_MINB) pack_reads_kernel(const uint8_t* __restrict__ ascii,
                                                                  const int64_t* __restrict__ offsets,
                                                                  int64_t U, int row_words,
                                                                  uint32_t* __restrict__ packed,
                                                                  int32_t* __restrict__ len_out,
                                                                  int32_t* __restrict__ bad_count,
                                                                  int k, const int32_t* __restrict__ segment,
                                                                  uint64_t* __restrict__ prefix_key,
                                                                  uint64_t* __restrict__ suffix_key) {
    __shared__ uint32_t bits[kPackPieces + 3];
    __shared__ int64_t s_span[2];
    const int quads = row_words >> 2;                // 16-byte groups per row
    const int64_t n_slots = U * quads;
    const int64_t slot = (int64_t)blockIdx.x * kPackThreads + threadIdx.x;
    const bool active = slot < n_slots;
    const int64_t total_bytes = offsets[U];          // issued first: only the tail check needs it
    int64_t u = 0, o0 = 0;
    int q = 0, len = 0;
    if (active) {
        // (u, q) = divmod(slot, quads): one 64-bit division per CTA (uniform), a 32-bit one per thread
        const int64_t slot0 = (int64_t)blockIdx.x * kPackThreads;
        const int64_t u0 = slot0 / quads;
        const unsigned local = (unsigned)(slot0 - u0 * quads) + threadIdx.x;
        const unsigned du = local / (unsigned)quads;
        u = u0 + du;
        q = (int)(local - du * (unsigned)quads);
        o0 = offsets[u];
        len = (int)(offsets[u + 1] - o0);
        if (q == 0) len_out[u] = len;
    }
    const int nvalid = len - 64 * q;                 // bases this thread holds (<= 0: none)
    const int64_t my_begin = o0 + min(64 * q, len);
    if (threadIdx.x == 0) s_span[0] = max((int64_t)0, my_begin - 32) & ~(int64_t)15;
    if (active && (threadIdx.x == kPackThreads - 1 || slot == n_slots - 1)) s_span[1] = o0 + min(64 * (q + 1), len);
    __syncthreads();
    const int64_t span_base = s_span[0];
    const int n_pieces = (int)((s_span[1] - span_base + 15) >> 4);
    uint32_t bad = 0;
    // all of the thread's 16-byte loads are issued before the first conversion (memory-level parallelism: a CTA's
    // span is at most kPackPieces pieces = kPackIters rounds)
    constexpr int kPre = OVL_PACK_PRELOADS < kPackIters ? OVL_PACK_PRELOADS : kPackIters;   // rounds whose load is issued up front
    uint4 v[kPre];
#pragma unroll
    for (int it = 0; it < kPre; ++it) {
        const int p = it * kPackThreads + threadIdx.x;
        v[it] = p < n_pieces ? __ldg(reinterpret_cast<const uint4*>(ascii + span_base + 16 * (int64_t)p)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int it = 0; it < kPackIters; ++it) {
        const int p = it * kPackThreads + threadIdx.x;
        if (p >= n_pieces + 3) break;
        uint32_t w = 0;
        if (p < n_pieces) {
            const int64_t g = span_base + 16 * (int64_t)p;
            const uint4 vv = it < kPre ? v[it < kPre ? it : 0] : __ldg(reinterpret_cast<const uint4*>(ascii + g));   // the last round holds at most 8 pieces
            uint32_t x[4] = {vv.x, vv.y, vv.z, vv.w};
            if (g + 16 > total_bytes) {              // the slack after the last read is not input: treat it as 'A'
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    int64_t nv = total_bytes - (g + 4 * i);
                    uint32_t keep = nv >= 4 ? 0xffffffffu : (nv <= 0 ? 0u : ((1u << (8 * (int)nv)) - 1u));
                    x[i] = (x[i] & keep) | (0x41414141u & ~keep);
                }
            }
            const uint32_t m0 = codes4(x[0], bad), m1 = codes4(x[1], bad), m2 = codes4(x[2], bad), m3 = codes4(x[3], bad);
            w = __byte_perm(__byte_perm(m0, m1, 0x0073), __byte_perm(m2, m3, 0x0073), 0x5410);
        }
        OVL_CHECK(p >= 0 && p < kPackPieces + 3);
        bits[p] = w;                                 // three zero words of padding after the stream
    }
    if (bad) atomicAdd(bad_count, 1);
    __syncthreads();
    if (!active) return;
    uint32_t o[4] = {0u, 0u, 0u, 0u};
    if (nvalid > 0) {
        const int bit0 = 2 * (int)(my_begin - span_base);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int nv = nvalid - 16 * w;
            if (nv > 0) {
                const int bit = bit0 + 32 * w;
                OVL_CHECK(bit >= 0 && (bit >> 5) + 1 < kPackPieces + 3);
                uint32_t x = __funnelshift_r(bits[bit >> 5], bits[(bit >> 5) + 1], bit & 31);
                if (nv < 16) x &= (1u << (2 * nv)) - 1u;
                o[w] = x;
            }
        }
    }
    reinterpret_cast<uint4*>(packed)[slot] = make_uint4(o[0], o[1], o[2], o[3]);
    if (KEYS) {
        const uint64_t mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
        const uint64_t tag = segment != nullptr ? (uint64_t)(uint32_t)segment[u] << (2 * k) : 0ull;
        if (q == 0)                                                       // overlapGraphs.py:33-37: prefix = read[:k]
            prefix_key[u] = len >= k ? (((((uint64_t)o[1] << 32) | o[0]) & mask) | tag) : kInvalidKey;
        if (q == (len > 0 ? (len - 1) >> 6 : 0)) {                        // :44-47: suffix = read[-k:]
            uint64_t sk = kInvalidKey;
            if (len >= k) {
                const int bit = 2 * (int)(o0 + len - k - span_base);      // >= 0: the span starts 32 bases before my slot
                const int i = bit >> 5, sh = bit & 31;
                OVL_CHECK(bit >= 0 && i + 2 < kPackPieces + 3);
                const uint32_t a = bits[i], b = bits[i + 1], c = bits[i + 2];
                sk = ((((uint64_t)__funnelshift_r(b, c, sh) << 32) | __funnelshift_r(a, b, sh)) & mask) | tag;
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
