/*
 * ovl.h -- C ABI of libovl_b200.so: the B200 (sm_100a) overlap-detection hot path of
 * roiteichman/Genome-Assembly-Using-Overlap-Graphs.
 *
 * The reference has no FFI layer: the boundary is two Python callables,
 *   aligners.overlap_alignment(s, t, match_score=10, mismatch=-1, indel=-2**31)   aligners.py:6-82
 *   overlapGraphs.construct_overlap_graph_nx_k(reads, k=5)                        overlapGraphs.py:5-61
 * The Python drop-ins of those two functions (package genome-assembly-using-overlap-graphs_b200)
 * bind the entry points below with ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns 0 on success, < 0 on error; ovl_last_error() gives the message
 *     (thread local).  No exceptions cross the boundary, no ownership is transferred.
 *   - every data pointer is a caller-owned DEVICE pointer unless its name starts with h_.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous w.r.t. the host.
 *   - scratch memory comes from the caller: query with the *_workspace_bytes() functions.
 *   - one context per device; a context is not thread safe.
 */
#ifndef OVL_H
#define OVL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVL_OK 0
#define OVL_E_CUDA (-1)      /* a CUDA runtime call or kernel launch failed */
#define OVL_E_ARG (-2)       /* invalid argument */
#define OVL_E_UNSUPPORTED (-3) /* valid for the reference, outside what the kernels implement */

#define OVL_MAX_K 32          /* k-mer keys are 2k-bit integers in a uint64; larger k: hashed keys + verify */
#define OVL_MAX_READ_LEN 2432 /* longest read the register-wavefront DP covers (32 lanes x 76 columns) */
#define OVL_MAX_LONG_READ_LEN 16384 /* longest read of the CTA-per-pair anti-diagonal DP (diagonals in shared memory) */

typedef struct ovl_ctx ovl_ctx;

const char *ovl_last_error(void);
int ovl_version(void);

int ovl_ctx_create(int device, ovl_ctx **out);
int ovl_ctx_destroy(ovl_ctx *ctx);
int ovl_ctx_sm_count(const ovl_ctx *ctx);
/* number of kernels launched through this context so far (measurement bookkeeping) */
int64_t ovl_ctx_launch_count(const ovl_ctx *ctx);

/* words per packed row for reads up to max_len bases: ceil(max_len/16) rounded up to a
 * multiple of 4 so that every row is 16-byte aligned. */
int32_t ovl_row_words(int32_t max_len);

/* K0: ASCII reads -> 2-bit packed rows.  Stands in for the reference's Python str reads
 * (the argument of overlapGraphs.py:5) in device form.
 *   ascii      U reads concatenated; 16-byte aligned, >= 32 bytes of slack after the end
 *   offsets    int64[U+1] byte offsets of the reads inside ascii
 *   packed     uint32[U * row_words] out; base i of a read at bits 2*(i%16) of word i/16
 *   len        int32[U] out
 *   bad_count  int32[1] in/out (caller zeroes): number of 16-base words holding a byte
 *              other than A, C, G, T.  The callers raise on bad_count != 0. */
int ovl_pack_reads(ovl_ctx *ctx, const uint8_t *ascii, const int64_t *offsets, int64_t U,
                   int32_t row_words, uint32_t *packed, int32_t *len, int32_t *bad_count,
                   void *stream);

/* K1: prefix / suffix k-mer keys, overlapGraphs.py:33-37 (read[:k]) and :44-47 (read[-k:]).
 * Reads shorter than k can match nothing but themselves: their keys are placeholders and the
 * index / join skip them by length. */
/* segment (optional, int32[U]): the read set each read belongs to.  The tag is stored in the key
 * bits above the k-mer (needs 2k + bits(segment) <= 64), so one index and one join serve many
 * independent read sets -- the parameter sweep of experiments.py:451-539 as ONE job -- and reads of
 * different sets never pair. */
int ovl_kmer_keys(ovl_ctx *ctx, const uint32_t *packed, int32_t row_words, const int32_t *len,
                  int64_t U, int32_t k, const int32_t *segment, uint64_t *prefix_key,
                  uint64_t *suffix_key, void *stream);

/* k > OVL_MAX_K: 64-bit hashes of the prefix / suffix k-mers instead of the k-mers themselves.  Build
 * the index on them with ovl_index_build(key_bits = 64) and use the *_verify join below, which
 * compares the actual k-mers of every hash match (exact result, same candidate order). */
int ovl_kmer_hashes(ovl_ctx *ctx, const uint32_t *packed, int32_t row_words, const int32_t *len,
                    int64_t U, int32_t k, uint64_t *prefix_hash, uint64_t *suffix_hash, void *stream);

/* K0 + K1 in one pass: the thread that packs a read's first / last bases also emits its prefix /
 * suffix k-mer key (1 <= k <= OVL_MAX_K; segment as in ovl_kmer_keys). */
int ovl_pack_reads_keys(ovl_ctx *ctx, const uint8_t *ascii, const int64_t *offsets, int64_t U,
                        int32_t row_words, int32_t k, const int32_t *segment, uint32_t *packed,
                        int32_t *len, int32_t *bad_count, uint64_t *prefix_key, uint64_t *suffix_key,
                        void *stream);

/* K2: the prefix index, overlapGraphs.py:30-40, as a stable sort of (prefix_key, uid):
 * sorted_key / sorted_uid hold the *n_indexed reads with len >= k, keys ascending and uids
 * ascending inside equal keys (the reference's bucket-append order).  LSD radix sort with digits of
 * up to 10 bits: one pass for k <= 5, two for k <= 10.
 * Optional outputs (NULL to skip):
 *   table   int32[2^table_bits + 1], a direct-address bucket table over the top table_bits bits of the
 *           key: table[t] = first sorted position whose key >> (key_bits - table_bits) is >= t.  With
 *           table_bits == key_bits (4^k <= 2^22) a bucket is table[key] .. table[key + 1] -- the
 *           dict lookup of overlapGraphs.py:49 without any search.  ovl_index_table_bits gives the
 *           size this library picks for (U, key_bits).
 *   pos_of  int32[U]: sorted position of every indexed read (its own slot in its bucket).
 *   copies / sorted_copies (both or neither): int32[U] multiplicity of every read in, the same along
 *           the sorted index out (sorted_copies[i] = copies[sorted_uid[i]], 0 for i >= *n_indexed) -- what
 *           ovl_join_count scans when reads have copies. */
size_t ovl_index_workspace_bytes(int64_t U);
int32_t ovl_index_table_bits(int64_t U, int32_t key_bits);
/* key_bits: number of significant key bits to sort on (2k + segment-tag bits); 0 means 2k. */
int ovl_index_build(ovl_ctx *ctx, const uint64_t *prefix_key, const int32_t *len, int64_t U, int32_t k,
                    int32_t key_bits, uint64_t *sorted_key, uint32_t *sorted_uid, int64_t *n_indexed,
                    int32_t *table, int32_t table_bits, int32_t *pos_of,
                    const int32_t *copies, int32_t *sorted_copies,
                    void *workspace, size_t workspace_bytes, void *stream);

/* K3: candidate generation, overlapGraphs.py:43-52, for all source reads a in [0, U).
 * ovl_join_count writes, per a, the bucket start, a's own rank inside the bucket (-1 if it
 * is not in it) and the exclusive scan pair_off[U+1] of the candidate counts; pair_off[U] is the
 * number of pairs.  table / pos_of (from ovl_index_build) are optional: without them the bucket
 * and the own slot are found by binary search.
 * With duplicate reads (copies != NULL; overlapGraphs.py:55-60 expands pair (a, b) to
 * copies[a] * copies[b] edges) it also writes cum[U+1], the exclusive scan of copies along the sorted
 * index (read from sorted_copies when ovl_index_build produced it, else gathered), and edge_base[U+1], the exclusive scan of the per-source edge counts: together with
 * bucket_lo / self_rank they give the first edge row of ANY pair in O(1), so no per-pair offset array
 * is ever built (ovl_overlap_dp_edges_join).
 * ovl_join_fill then writes pairs [p_begin, p_begin+p_count), ordered by (a, b) ascending.
 * total_hint = the total number of pairs (pair_off[U], which the host has read anyway): it
 * only selects between the thread-per-pair and the warp-per-source fill kernels. */
size_t ovl_join_workspace_bytes(int64_t n_sources);
int ovl_join_count(ovl_ctx *ctx, const uint64_t *suffix_key, const uint64_t *prefix_key,
                   const int32_t *len, int32_t k, int64_t U, const uint64_t *sorted_key,
                   const uint32_t *sorted_uid, const int64_t *n_indexed, const int32_t *table,
                   int32_t table_bits, int32_t key_bits, const int32_t *pos_of, const int32_t *copies,
                   const int32_t *sorted_copies, int64_t *cum, int32_t *bucket_lo, int32_t *self_rank,
                   int64_t *pair_off, int64_t *edge_base, void *workspace, size_t workspace_bytes,
                   void *stream);
/* The scalars the host needs before it can size anything, in ONE block of ovl_totals_len() int64
 * words written by one small kernel (totals may be page-locked host memory: the host then needs a
 * stream/event wait and no copy): [0] pairs, [1] edge rows, [2] *bad_count, [3],[4] this rank's
 * slice [p_begin, p_end) of the pair list (rank of world equal slices), [5] *n_indexed,
 * [8 .. 8+64] pair indices cutting the slice into 64 equal parts, [73 .. 73+64] the first edge row
 * of each of those pairs (D2H chunk boundaries, shard boundaries). */
int32_t ovl_totals_len(void);
int ovl_join_finalize(ovl_ctx *ctx, const int64_t *pair_off, const int64_t *edge_base,
                      const int32_t *bucket_lo, const int32_t *self_rank, const int64_t *cum,
                      const int32_t *copies, int64_t U, const int32_t *bad_count,
                      const int64_t *n_indexed, int32_t rank, int32_t world, int64_t *totals,
                      void *stream);
int ovl_join_fill(ovl_ctx *ctx, const int64_t *pair_off, int64_t a_begin, int64_t a_end,
                  const int32_t *bucket_lo, const int32_t *self_rank, const uint32_t *sorted_uid,
                  int64_t p_begin, int64_t p_count, int64_t total_hint, int32_t *pair_a,
                  int32_t *pair_b, void *stream);
/* The same join on hashed keys (k > OVL_MAX_K): every hash match is verified base by base. */
int ovl_join_count_verify(ovl_ctx *ctx, const uint32_t *packed, int32_t row_words, const int32_t *len,
                          int32_t k, const uint64_t *suffix_hash, int64_t a_begin, int64_t a_end,
                          const uint64_t *sorted_hash, const uint32_t *sorted_uid,
                          const int64_t *n_indexed, int64_t *pair_off, void *workspace,
                          size_t workspace_bytes, void *stream);
int ovl_join_fill_verify(ovl_ctx *ctx, const uint32_t *packed, int32_t row_words, const int32_t *len,
                         int32_t k, const uint64_t *suffix_hash, int64_t a_begin, int64_t a_end,
                         const uint64_t *sorted_hash, const uint32_t *sorted_uid,
                         const int64_t *n_indexed, const int64_t *pair_off, int64_t p_begin,
                         int64_t p_count, int32_t *pair_a, int32_t *pair_b, void *stream);
/* k == 0 (overlapGraphs.py:49): all ordered pairs a != b, a in [a_begin, ...); pair index p
 * counts from a_begin: a = a_begin + p / (U-1). */
int ovl_all_pairs_fill(ovl_ctx *ctx, int64_t U, int64_t a_begin, int64_t p_begin, int64_t p_count,
                       int32_t *pair_a, int32_t *pair_b, void *stream);

/* K0-K3 in one call (2-bit reads, 1 <= k <= OVL_MAX_K): pack + keys, index + table, join count,
 * totals -- a dozen kernels launched back to back on `stream`, all arrays carved out of one
 * caller-owned arena.  ovl_candidates_layout fills the byte offsets (and the sizes it chose);
 * arena must be 256-byte aligned and lay->total_bytes long.  After the call (and a wait on the
 * stream) `totals` holds the block described at ovl_join_finalize (pos_of is not built: lay->pos_of is 0); the caller then sizes the pair
 * list and calls ovl_join_fill / ovl_overlap_dp_edges[_join] on the arrays inside the arena. */
typedef struct ovl_cand_layout {
    size_t packed, len, bad, n_indexed, prefix_key, suffix_key, sorted_key, sorted_uid, table, pos_of,
           bucket_lo, self_rank, pair_off, edge_base, cum, sorted_copies, scratch, scratch_bytes, total_bytes;
    int32_t row_words, key_bits, table_bits, has_copies;
} ovl_cand_layout;
int ovl_candidates_layout(int64_t U, int32_t max_len, int32_t k, int32_t n_segments, int32_t has_copies,
                          ovl_cand_layout *out);
int ovl_candidates_build(ovl_ctx *ctx, const uint8_t *ascii, const int64_t *offsets, int64_t U, int32_t k,
                         const int32_t *segments, const int32_t *copies, int32_t rank, int32_t world,
                         void *arena, const ovl_cand_layout *lay, int64_t *totals, void *stream);

/* K4/K5: the DP call site overlapGraphs.py:53, i.e. aligners.py:27-57 for every pair:
 *   score[p], end[p] = overlap_alignment(read[pair_a[p]], read[pair_b[p]], match, mismatch, indel)[3:5]
 * indel is int64 like the reference's Numba-typed default (-2**31 never wraps).
 * max_len = longest read in the batch (<= OVL_MAX_LONG_READ_LEN, <= 16*row_words); batches whose longest
 * read exceeds OVL_MAX_READ_LEN run the slower CTA-per-pair anti-diagonal kernel.
 * mode: 0 = choose, 1 = force the packed 16-bit kernel, 2 = force the 32-bit kernel.
 * group_lanes / cols_per_lane: 0 = choose, else force that instantiation (tests). */
int ovl_overlap_dp(ovl_ctx *ctx, const uint32_t *packed, int32_t row_words, const int32_t *len,
                   const int32_t *pair_a, const int32_t *pair_b, int64_t P, int32_t max_len,
                   int64_t match, int64_t mismatch, int64_t indel, int32_t *score, int32_t *end,
                   int32_t mode, int32_t group_lanes, int32_t cols_per_lane, void *stream);
/* K4/K5 with K6 fused into the epilogue: instead of score/end the kernel writes the pair's
 * copy_a x copy_b edge rows (overlapGraphs.py:55-60) itself.  copies == NULL: every read occurs
 * once, edges[p] = (a, b, score, end).  Otherwise edge_off[P+1] comes from ovl_expand_count and
 * edges has edge_off[P] rows. */
int ovl_overlap_dp_edges(ovl_ctx *ctx, const uint32_t *packed, int32_t row_words, const int32_t *len,
                         const int32_t *pair_a, const int32_t *pair_b, int64_t P, int32_t max_len,
                         int64_t match, int64_t mismatch, int64_t indel, const int32_t *copies,
                         const int64_t *node_off, const int64_t *edge_off, int32_t *edges,
                         void *stream);
/* The same with the edge-row offsets taken from the join index (ovl_join_count with copies) instead of
 * a per-pair edge_off array: pair_a[0] is global pair p_begin, edges[0] is global edge row e_begin. */
int ovl_overlap_dp_edges_join(ovl_ctx *ctx, const uint32_t *packed, int32_t row_words, const int32_t *len,
                              const int32_t *pair_a, const int32_t *pair_b, int64_t P, int32_t max_len,
                              int64_t match, int64_t mismatch, int64_t indel, const int32_t *copies,
                              const int64_t *node_off, const int64_t *pair_off, const int64_t *edge_base,
                              const int32_t *bucket_lo, const int32_t *self_rank, const int64_t *cum,
                              int64_t p_begin, int64_t e_begin, int32_t *edges, void *stream);
/* which kernel ovl_overlap_dp would pick: out[0]=mode (1 packed, 2 int32, 3 long-read kernel), out[1]=lanes,
 * out[2]=columns per lane.  Returns OVL_E_UNSUPPORTED when nothing fits. */
int ovl_overlap_dp_plan(int32_t max_len, int64_t match, int64_t mismatch, int64_t indel,
                        int32_t mode, int32_t out[3]);

/* Byte-coded reads: read sets with more than four distinct symbols cannot be 2-bit packed (the
 * reference compares arbitrary characters, aligners.py:35).  They are kept as padded byte rows
 * (row_words*4 bytes per read) and take the general route: hashed keys + byte-wise verified join and
 * the CTA-per-pair anti-diagonal DP.  Same results, lower throughput.  ovl_overlap_dp8 writes
 * score/end when edges == NULL, else the fused edge rows (as ovl_overlap_dp_edges). */
int ovl_pack_bytes(ovl_ctx *ctx, const uint8_t *ascii, const int64_t *offsets, int64_t U,
                   int32_t row_words, uint8_t *rows, int32_t *len, void *stream);
int ovl_kmer_hashes8(ovl_ctx *ctx, const uint8_t *rows, int32_t row_words, const int32_t *len,
                     int64_t U, int32_t k, uint64_t *prefix_hash, uint64_t *suffix_hash, void *stream);
int ovl_join_count_verify8(ovl_ctx *ctx, const uint8_t *rows, int32_t row_words, const int32_t *len,
                           int32_t k, const uint64_t *suffix_hash, int64_t a_begin, int64_t a_end,
                           const uint64_t *sorted_hash, const uint32_t *sorted_uid,
                           const int64_t *n_indexed, int64_t *pair_off, void *workspace,
                           size_t workspace_bytes, void *stream);
int ovl_join_fill_verify8(ovl_ctx *ctx, const uint8_t *rows, int32_t row_words, const int32_t *len,
                          int32_t k, const uint64_t *suffix_hash, int64_t a_begin, int64_t a_end,
                          const uint64_t *sorted_hash, const uint32_t *sorted_uid,
                          const int64_t *n_indexed, const int64_t *pair_off, int64_t p_begin,
                          int64_t p_count, int32_t *pair_a, int32_t *pair_b, void *stream);
int ovl_overlap_dp8(ovl_ctx *ctx, const uint8_t *rows, int32_t row_words, const int32_t *len,
                    const int32_t *pair_a, const int32_t *pair_b, int64_t P, int32_t max_len,
                    int64_t match, int64_t mismatch, int64_t indel, int32_t *score, int32_t *end,
                    const int32_t *copies, const int64_t *node_off, const int64_t *edge_off,
                    int32_t *edges, void *stream);

/* K6: edge expansion, overlapGraphs.py:55-60.  Edge row = int32[4] (node_a, node_b, weight,
 * end_position); node id = node_off[uid] + copy; order = pair order, copy_a, copy_b.
 * ovl_expand_count writes the exclusive scan edge_off[P+1] of copies[a]*copies[b]. */
size_t ovl_expand_workspace_bytes(int64_t P);
int ovl_expand_count(ovl_ctx *ctx, const int32_t *pair_a, const int32_t *pair_b,
                     const int32_t *copies, int64_t P, int64_t *edge_off, void *workspace,
                     size_t workspace_bytes, void *stream);
int ovl_expand_fill(ovl_ctx *ctx, const int64_t *edge_off, int64_t P, const int32_t *pair_a,
                    const int32_t *pair_b, const int32_t *score, const int32_t *end,
                    const int32_t *copies, const int64_t *node_off, int64_t e_begin,
                    int64_t e_count, int32_t *edges, void *stream);
/* every read occurs once (copies == 1 everywhere): edges[p] = (a, b, score, end) */
int ovl_expand_unit(ovl_ctx *ctx, const int32_t *pair_a, const int32_t *pair_b,
                    const int32_t *score, const int32_t *end, int64_t P, int32_t *edges,
                    void *stream);

/* The `if score > 0` of the all-pairs builders (overlapGraphs.py:225, :347): order-preserving
 * compaction of edge rows with weight >= min_weight.  ovl_filter_count writes the exclusive scan
 * keep_off[E+1] of the keep flags (keep_off[E] = rows kept); ovl_filter_fill writes them. */
size_t ovl_filter_workspace_bytes(int64_t E);
int ovl_filter_count(ovl_ctx *ctx, const int32_t *edges, int64_t E, int32_t min_weight,
                     int64_t *keep_off, void *workspace, size_t workspace_bytes, void *stream);
int ovl_filter_fill(ovl_ctx *ctx, const int32_t *edges, const int64_t *keep_off, int64_t E,
                    int32_t min_weight, int32_t *out, void *stream);

/* K7: one pair with traceback: everything aligners.py:27-76 computes, for the single-pair
 * drop-in.  s, t are int32 code points (any alphabet); arithmetic is the reference's (int64
 * candidates, int32 storage).  result[0..2] = (best_score, alignment_end_position, n_ops);
 * ops[0..n_ops) is the traceback from the end backwards: 0 diagonal, 1 up (gap in t),
 * 2 left (gap in s); ops needs n+m bytes. */
size_t ovl_align_pair_workspace_bytes(int32_t n, int32_t m);
int ovl_align_pair(ovl_ctx *ctx, const int32_t *s, int32_t n, const int32_t *t, int32_t m,
                   int64_t match, int64_t mismatch, int64_t indel, void *workspace,
                   size_t workspace_bytes, int32_t *result, uint8_t *ops, void *stream);

/* K8: Smith-Waterman local alignment with traceback = aligners.local_alignment, aligners.py:85-167
 * (the aligner the evaluation uses to map reads / contigs to the genome, performanceMeasures.py:219).
 * result[0..4] = (best_score, start_pos, end_pos, n_ops, best_i); ops[0..n_ops) is the traceback
 * from the best cell backwards: 1 diagonal, 2 up (gap in reference), 3 left (gap in query). */
size_t ovl_local_align_workspace_bytes(int32_t n, int32_t m);
int ovl_local_align(ovl_ctx *ctx, const int32_t *query, int32_t n, const int32_t *reference, int32_t m,
                    int64_t match, int64_t mismatch, int64_t indel, void *workspace,
                    size_t workspace_bytes, int32_t *result, uint8_t *ops, void *stream);

/* K8 batch: many queries against windows of ONE reference in a single launch, one CTA per query
 * (performanceMeasures.py:219-221 calls align_read_or_contig_to_reference once per contig, always with
 * the same genome).  Per query x (all arrays on the device): symbols queries[q_off[x] .. q_off[x+1]),
 * at most OVL_LOCAL_BATCH_MAX_QUERY of them (max_query_len = the longest); reference window
 * [ref_start[x], ref_start[x] + ref_len[x]); tb_off[x] = byte offset of its traceback area inside tb,
 * (n + 1) * (n + m + 2) bytes; ops_off[x] = byte offset of its op list inside ops, n + m + 1 bytes;
 * results + 8 * x receives the same five words as ovl_local_align.  Results equal one
 * ovl_local_align call per query. */
#define OVL_LOCAL_BATCH_MAX_QUERY 1024
int ovl_local_align_batch(ovl_ctx *ctx, const int32_t *queries, const int64_t *q_off, int32_t n_queries,
                          int32_t max_query_len, const int32_t *reference, const int32_t *ref_start,
                          const int32_t *ref_len, int64_t match, int64_t mismatch, int64_t indel,
                          uint8_t *tb, const int64_t *tb_off, int32_t *results, uint8_t *ops,
                          const int64_t *ops_off, void *stream);

/* Seeded read simulator (the input generators generateErrorFreeReads.py:22-52 and
 * generateErrorProneReads.py:4-45 as one counter-based stream): n_reads reads of read_len bases with
 * uniform starts on the LINEAR genome (A/C/G/T ASCII, genome_len < 2^32; reads are truncated at its
 * end), every base replaced with probability error_thr / 2^32 by one of the three other letters.
 * Writes offsets[n_reads + 1] and the reads back to back into ascii (capacity n_reads * read_len + 64,
 * 16-byte aligned: directly the input of ovl_pack_reads / ovl_candidates_build).  Same bytes as
 * synth.simulate_reads_counter() for the same seed. */
size_t ovl_simulate_workspace_bytes(int64_t n_reads);
int ovl_simulate_reads(ovl_ctx *ctx, const uint8_t *genome, int64_t genome_len, int64_t n_reads,
                       int32_t read_len, uint32_t error_thr, uint64_t seed, int64_t *offsets,
                       uint8_t *ascii, void *workspace, size_t workspace_bytes, void *stream);

/* Pre-pass for the weakest-edge cycle removal (overlapGraphs.py:106-130): peel sinks off the directed graph
 * given as an edge list (src[e] -> dst[e], node ids < n_nodes, no self loops) until none is left.
 * state[v] (int32[n_nodes], out) = 0 if v survives -- it reaches a cycle or lies on one -- else the
 * round (>= 1) in which it became a sink.  Peeled nodes cannot be on any cycle nx.find_cycle reports and
 * leaving them out does not change which cycle it reports, so the host loop runs on the survivors only
 * and removes the same edges in the same order.  Synchronises the stream (the number of rounds is
 * data dependent); *h_rounds (host, optional) receives the rounds launched. */
size_t ovl_trim_workspace_bytes(int64_t n_nodes);
int ovl_trim_sinks(ovl_ctx *ctx, const int32_t *src, const int32_t *dst, int64_t E, int64_t n_nodes,
                   int32_t *state, void *workspace, size_t workspace_bytes, int32_t *h_rounds,
                   void *stream);

/* Order-sensitive fingerprint of an edge list: adds, into *accum (device u64, zeroed by the caller),
 * the sum over rows of mix(first_row + i, row i) mod 2^64.  Shards hashed with their global row
 * offset add up to the fingerprint of the whole list; any misplaced or reordered row changes it.
 * Self-check of the multi-GPU exchange paths (no counterpart in the reference). */
int ovl_edge_list_hash(ovl_ctx *ctx, const int32_t *edges, int64_t E, int64_t first_row,
                       uint64_t *accum, void *stream);

/* Roofline denominator for the DP: runs a dependency-free instruction stream on every SM and
 * returns lane-operations per second (1e9/s).  kind: 0 IADD3, 1 IMAD, 2 VIMNMX.S32,
 * 3 VIADDMNMX.S16x2, 4 the DP inner-loop mix (PRMT, IMAD, 2x VIADDMNMX.S16x2), 5 PRMT, 6 LOP3,
 * 7 LOP3 + IMAD on independent chains (do the ALU and FMA pipes issue side by side?),
 * 8 VIMNMX3 + IMAD with all-distinct register operands, 9 one form-1 DP column per chain,
 * 10 one form-2 DP column per chain, 11 a 2 ALU + 2 IMAD column, 12-16 pipe-pairing experiments,
 * 17 __vminu2 (2-input packed min: compiles to VIMNMX3.U16x2 with a repeated operand), 18 VIMNMX3.U16x2,
 * 20 VIADDMNMX.U16x2, 21 VIADDMNMX.U16x2 with an immediate addend (two register sources), 22 a form-1 column
 * with immediate gap costs, 23 IMAD with an immediate multiplier, 24 PRMT with a repeated source, 25 LOP3 with
 * an immediate, 26 HMNMX2 (fp16x2 min), 27-30 HMNMX2 next to LOP3 / IMAD / VIADDMNMX.U16x2 / PRMT.
 * Synchronises the device.  h_gops receives giga lane-instructions per second. */
int ovl_int_peak_probe(ovl_ctx *ctx, int32_t kind, int32_t iters, double *h_gops, double *h_ms);

#ifdef __cplusplus
}
#endif
#endif /* OVL_H */
