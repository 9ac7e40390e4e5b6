// Implicit edge offsets of the k-mer join.
//
// With duplicate reads, candidate pair p = (a, b) expands to copies[a] * copies[b] edge rows
// (overlapGraphs.py:55-60) and the rows of all pairs are laid out back to back in pair order.  The
// row offset of a pair is NOT materialised per pair (that would be an 8-byte entry and a scan over
// ~1e9 pairs): it follows from four read-sized arrays the join already has --
//     edge_off(p) = edge_base[a] + copies[a] * (cum[pos] - cum[lo[a]] - (pos is past a's own slot ? copies[a] : 0))
// where [lo[a], hi) is a's bucket in the sorted prefix index, pos the sorted position of b, cum the
// exclusive scan of copies over the sorted index, and edge_base the exclusive scan of the per-source
// edge counts.  The DP epilogue and the host-side slicing (shards, D2H chunks) both use it.
#pragma once
#include "common.cuh"

namespace ovl {

struct JoinEdgeIndex {
    const int64_t* pair_off;    // [U+1] exclusive scan of per-source candidate counts; null = not in use
    const int64_t* edge_base;   // [U+1] exclusive scan of per-source edge counts
    const int32_t* lo;          // [U]   first sorted position of the source's bucket
    const int32_t* self_rank;   // [U]   the source's own rank inside its bucket, or -1
    const int64_t* cum;         // [n_indexed+1] exclusive scan of copies[sorted_uid[.]]
    int64_t p_begin;            // global pair index of the launch's pair 0
    int64_t e_begin;            // global edge index of the launch's output row 0
};

__device__ __forceinline__ int64_t join_edge_offset(const JoinEdgeIndex& jx, const int32_t* __restrict__ copies,
                                                    int64_t p_global, int32_t a) {
    const int64_t r = p_global - jx.pair_off[a];
    const int32_t sr = jx.self_rank[a];
    const int64_t lo = jx.lo[a];
    const bool past_self = sr >= 0 && r >= sr;
    const int64_t pos = lo + r + (past_self ? 1 : 0);
    const int64_t ca = copies[a];
    return jx.edge_base[a] + ca * (jx.cum[pos] - jx.cum[lo] - (past_self ? ca : 0));
}

}  // namespace ovl
