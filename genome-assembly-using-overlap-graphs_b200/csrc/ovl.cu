// libovl_b200.so -- C ABI over the sm_100a overlap-detection kernels (see include/ovl.h).
#include "../../include/ovl.h"

#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>

#include "common.cuh"
#include "scan.cuh"
#include "index.cuh"
#include "kmer.cuh"
#include "dp.cuh"
#include "probe.cuh"
#include "align.cuh"
#include "check.cuh"
#include "simulate.cuh"
#include "trim.cuh"

using namespace ovl;

struct ovl_ctx {
    int device;
    int sm_count;
    uint32_t* probe_sink;
    void* scratch;               // 256 bytes of device memory for tiny hand-offs between kernels of one call
    long long launches;          // kernels launched through this context (bench bookkeeping)
};

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return fail(OVL_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                \
    } while (0)

#define LAUNCH_CHECK(name)                                                                   \
    do {                                                                                     \
        ctx->launches += 1;                                                                  \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return fail(OVL_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

// Every entry point that takes a context runs on the context's device and leaves the calling
// thread's current device as it found it (PyTorch owns that state).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(const ovl_ctx* ctx) {
        if (ctx && cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device)
            switched = cudaSetDevice(ctx->device) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ON_CTX_DEVICE(ctx) DeviceGuard device_guard__(ctx)

static inline unsigned grid_for(int64_t n, int per_block) {
    int64_t g = (n + per_block - 1) / per_block;
    return (unsigned)(g > 0 ? g : 1);
}

extern "C" {

const char* ovl_last_error(void) { return g_err; }
int ovl_version(void) { return 100; }

int ovl_ctx_create(int device, ovl_ctx** out) {
    if (!out) return fail(OVL_E_ARG, "ovl_ctx_create: out is null");
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(OVL_E_ARG, "ovl_ctx_create: no CUDA device %d (have %d)", device, count);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));       // does not change the current device
    if (prop.major != 10)
        return fail(OVL_E_UNSUPPORTED, "ovl_ctx_create: device %d is sm_%d%d; this library holds sm_100a code only",
                    device, prop.major, prop.minor);
    ovl_ctx* c = new ovl_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->probe_sink = nullptr;
    c->scratch = nullptr;
    c->launches = 0;
    *out = c;
    return OVL_OK;
}

int ovl_ctx_destroy(ovl_ctx* ctx) {
    if (!ctx) return OVL_OK;
    ON_CTX_DEVICE(ctx);
    if (ctx->probe_sink) cudaFree(ctx->probe_sink);
    if (ctx->scratch) cudaFree(ctx->scratch);
    delete ctx;
    return OVL_OK;
}

int ovl_ctx_sm_count(const ovl_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
int64_t ovl_ctx_launch_count(const ovl_ctx* ctx) { return ctx ? (int64_t)ctx->launches : 0; }

int32_t ovl_row_words(int32_t max_len) {
    int32_t w = (max_len + 15) / 16;
    if (w < 1) w = 1;
    return (w + 3) & ~3;
}

// ---------------------------------------------------------------- K0 / K1
static int pack_launch(ovl_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int64_t U, int32_t row_words, int32_t k,
                       const int32_t* segment, uint32_t* packed, int32_t* len, int32_t* bad_count, uint64_t* prefix_key,
                       uint64_t* suffix_key, cudaStream_t st, const char* who) {
    if (!ctx || !ascii || !offsets || !packed || !len || !bad_count) return fail(OVL_E_ARG, "%s: null argument", who);
    if (U <= 0) return OVL_OK;
    if (row_words < 4 || (row_words & 3)) return fail(OVL_E_ARG, "%s: row_words must be a positive multiple of 4", who);
    if (((uintptr_t)ascii & 15) || ((uintptr_t)packed & 15)) return fail(OVL_E_ARG, "%s: ascii and packed must be 16-byte aligned", who);
    unsigned grid = grid_for(U * (row_words / 4), 256);
    if (prefix_key != nullptr) {
        if (k < 1 || k > OVL_MAX_K) return fail(OVL_E_UNSUPPORTED, "%s: k=%d outside 1..%d", who, k, OVL_MAX_K);
        if (segment && k > 31) return fail(OVL_E_UNSUPPORTED, "%s: segment tags need 2k < 64 (k=%d)", who, k);
        if (!suffix_key) return fail(OVL_E_ARG, "%s: prefix_key given without suffix_key", who);
        pack_reads_kernel<true><<<grid, 256, 0, st>>>(ascii, offsets, U, row_words, packed, len, bad_count, k, segment, prefix_key, suffix_key);
    } else {
        pack_reads_kernel<false><<<grid, 256, 0, st>>>(ascii, offsets, U, row_words, packed, len, bad_count, 0, nullptr, nullptr, nullptr);
    }
    LAUNCH_CHECK("pack_reads_kernel");
    return OVL_OK;
}

int ovl_pack_reads(ovl_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int64_t U, int32_t row_words,
                   uint32_t* packed, int32_t* len, int32_t* bad_count, void* stream) {
    ON_CTX_DEVICE(ctx);
    return pack_launch(ctx, ascii, offsets, U, row_words, 0, nullptr, packed, len, bad_count, nullptr, nullptr, (cudaStream_t)stream, "ovl_pack_reads");
}

int ovl_pack_reads_keys(ovl_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int64_t U, int32_t row_words, int32_t k,
                        const int32_t* segment, uint32_t* packed, int32_t* len, int32_t* bad_count, uint64_t* prefix_key,
                        uint64_t* suffix_key, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!prefix_key || !suffix_key) return fail(OVL_E_ARG, "ovl_pack_reads_keys: null key output");
    return pack_launch(ctx, ascii, offsets, U, row_words, k, segment, packed, len, bad_count, prefix_key, suffix_key, (cudaStream_t)stream,
                       "ovl_pack_reads_keys");
}

int ovl_kmer_keys(ovl_ctx* ctx, const uint32_t* packed, int32_t row_words, const int32_t* len, int64_t U, int32_t k,
                  const int32_t* segment, uint64_t* prefix_key, uint64_t* suffix_key, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !packed || !len || !prefix_key || !suffix_key) return fail(OVL_E_ARG, "ovl_kmer_keys: null argument");
    if (k < 1 || k > OVL_MAX_K) return fail(OVL_E_UNSUPPORTED, "ovl_kmer_keys: k=%d outside 1..%d", k, OVL_MAX_K);
    if (U <= 0) return OVL_OK;
    if (segment && k > 31) return fail(OVL_E_UNSUPPORTED, "ovl_kmer_keys: segment tags need 2k < 64 (k=%d)", k);
    kmer_keys_kernel<<<grid_for(U, 256), 256, 0, (cudaStream_t)stream>>>(packed, row_words, len, U, k, segment, prefix_key, suffix_key);
    LAUNCH_CHECK("kmer_keys_kernel");
    return OVL_OK;
}

int ovl_kmer_hashes(ovl_ctx* ctx, const uint32_t* packed, int32_t row_words, const int32_t* len, int64_t U, int32_t k,
                    uint64_t* prefix_key, uint64_t* suffix_key, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !packed || !len || !prefix_key || !suffix_key) return fail(OVL_E_ARG, "ovl_kmer_hashes: null argument");
    if (k < 1) return fail(OVL_E_ARG, "ovl_kmer_hashes: k must be positive");
    if (U <= 0) return OVL_OK;
    kmer_hash_kernel<<<grid_for(U, 256), 256, 0, (cudaStream_t)stream>>>(packed, row_words, len, U, k, prefix_key, suffix_key);
    LAUNCH_CHECK("kmer_hash_kernel");
    return OVL_OK;
}

// ---------------------------------------------------------------- K2
// workspace: [hist int32 2^D*C + 1][scan sums][tmp keys u64 U][tmp uids u32 U]     (C = CTAs of kSortChunk = 4,096 elements)
static inline int64_t sort_warps(int64_t U) { return (U + kSortChunk - 1) / kSortChunk; }
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// the digit totals of one sort pass (or the tile sums of a generic scan, whichever is larger)
static inline size_t index_sums_bytes(int64_t hn) {
    return align256(std::max(scan_workspace_bytes(hn, sizeof(int32_t)), (((size_t)1 << kSortMaxDigit) + 1) * sizeof(int32_t)));
}

size_t ovl_index_workspace_bytes(int64_t U) {
    if (U < 1) U = 1;
    int64_t W = sort_warps(U);
    int64_t hn = ((int64_t)1 << kSortMaxDigit) * W;
    size_t hist = align256((size_t)(hn + 1) * sizeof(int32_t));
    size_t sums = index_sums_bytes(hn);
    size_t keys = align256((size_t)U * sizeof(uint64_t));
    size_t uids = align256((size_t)U * sizeof(uint32_t));
    return hist + sums + keys + uids + 256;
}

int32_t ovl_index_table_bits(int64_t U, int32_t key_bits) {
    if (key_bits < 1) return 0;
    int lg = 0;
    while (((int64_t)1 << lg) < U) ++lg;
    int tb = std::min<int>(key_bits, kTableMaxBits);
    return std::min(tb, std::max(8, lg + 2));
}

int ovl_index_build(ovl_ctx* ctx, const uint64_t* prefix_key, const int32_t* len, int64_t U, int32_t k, int32_t key_bits, uint64_t* sorted_key,
                    uint32_t* sorted_uid, int64_t* n_indexed, int32_t* table, int32_t table_bits, int32_t* pos_of,
                    const int32_t* copies, int32_t* sorted_copies, void* workspace, size_t workspace_bytes, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !prefix_key || !len || !sorted_key || !sorted_uid || !n_indexed || !workspace) return fail(OVL_E_ARG, "ovl_index_build: null argument");
    if (k < 1) return fail(OVL_E_ARG, "ovl_index_build: k must be positive");
    if (k > OVL_MAX_K && key_bits != 64) return fail(OVL_E_ARG, "ovl_index_build: k=%d > %d needs hashed keys (key_bits = 64)", k, OVL_MAX_K);
    if (workspace_bytes < ovl_index_workspace_bytes(U)) return fail(OVL_E_ARG, "ovl_index_build: workspace too small");
    if (U > 0x7fffffffll) return fail(OVL_E_UNSUPPORTED, "ovl_index_build: more than 2^31 - 1 reads");
    if ((copies != nullptr) != (sorted_copies != nullptr)) return fail(OVL_E_ARG, "ovl_index_build: copies and sorted_copies go together");
    cudaStream_t st = (cudaStream_t)stream;
    if (key_bits <= 0) key_bits = 2 * k;                 // no segment tag above the k-mer
    if ((key_bits < 2 * k && k <= OVL_MAX_K) || key_bits > 64) return fail(OVL_E_ARG, "ovl_index_build: key_bits=%d outside [2k, 64]", key_bits);
    if (table && (table_bits < 1 || table_bits > key_bits || table_bits > kTableMaxBits))
        return fail(OVL_E_ARG, "ovl_index_build: table_bits=%d outside [1, min(key_bits, %d)]", table_bits, kTableMaxBits);
    if (U <= 0) {
        CUDA_TRY(cudaMemsetAsync(n_indexed, 0, sizeof(int64_t), st));
        if (table) CUDA_TRY(cudaMemsetAsync(table, 0, (((size_t)1 << table_bits) + 1) * sizeof(int32_t), st));
        return OVL_OK;
    }
    int64_t W = sort_warps(U);
    int64_t hn_max = ((int64_t)1 << kSortMaxDigit) * W;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int32_t* hist = (int32_t*)ws;                 ws += align256((size_t)(hn_max + 1) * sizeof(int32_t));
    int32_t* digit_total = (int32_t*)ws;          ws += index_sums_bytes(hn_max);
    uint64_t* tmp_key = (uint64_t*)ws;            ws += align256((size_t)U * sizeof(uint64_t));
    uint32_t* tmp_uid = (uint32_t*)ws;

    if (sorted_copies) CUDA_TRY(cudaMemsetAsync(sorted_copies, 0, (size_t)U * sizeof(int32_t), st));   // zero past the end of the index
    if (kSortStageBytes + 24 * 1024 > 48 * 1024) {      // the kernel also has ~22 KB of static shared memory
        static bool attr_set = false;               // per process: the attribute belongs to the function, not the context
        if (!attr_set) {
            CUDA_TRY(cudaFuncSetAttribute(sort_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortStageBytes));
            CUDA_TRY(cudaFuncSetAttribute(sort_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortStageBytes));
            attr_set = true;
        }
    }
    const int passes = sort_passes(key_bits);
    const int D = sort_digit_bits(key_bits);
    int nl = 0;
    // ping-pong so that the last pass lands in (sorted_key, sorted_uid)
    uint64_t* kbuf[2] = {sorted_key, tmp_key};
    uint32_t* ubuf[2] = {sorted_uid, tmp_uid};
    int dst = (passes & 1) ? 0 : 1;
    unsigned grid = (unsigned)W;                        // one CTA per kSortChunk-element chunk
    const uint64_t* src_key = prefix_key;
    const uint32_t* src_uid = nullptr;
    for (int p = 0; p < passes; ++p) {
        int shift = D * p;
        int bits = std::min(D, key_bits - shift);
        const bool last_pass = p == passes - 1;
        // a single pass over the whole key with a table as wide as the key: the scatter writes the table itself
        int32_t* table_direct = (table && passes == 1 && table_bits == key_bits) ? table : nullptr;
        int32_t* pos_out = last_pass ? pos_of : nullptr;
        int32_t* sc_out = last_pass ? sorted_copies : nullptr;
        if (p == 0) {
            sort_hist_kernel<true><<<grid, kSortThreads, 0, st>>>(src_key, len, k, nullptr, U, shift, bits, W, hist);
        } else {
            sort_hist_kernel<false><<<grid, kSortThreads, 0, st>>>(src_key, len, k, n_indexed, 0, shift, bits, W, hist);
        }
        LAUNCH_CHECK("sort_hist_kernel");
        sort_digit_scan_kernel<<<1u << bits, kScanThreads, 0, st>>>(hist, W, digit_total);
        LAUNCH_CHECK("sort_digit_scan_kernel");
        if (p == 0) {
            sort_scatter_kernel<true><<<grid, kSortThreads, kSortStageBytes, st>>>(src_key, nullptr, len, k, nullptr, U, shift, bits, W, hist, digit_total,
                                                                     table_direct, kbuf[dst], ubuf[dst], pos_out, copies, sc_out, n_indexed);
        } else {
            sort_scatter_kernel<false><<<grid, kSortThreads, kSortStageBytes, st>>>(src_key, src_uid, len, k, n_indexed, 0, shift, bits, W, hist, digit_total,
                                                                      table_direct, kbuf[dst], ubuf[dst], pos_out, copies, sc_out, nullptr);
        }
        LAUNCH_CHECK("sort_scatter_kernel");
        src_key = kbuf[dst];
        src_uid = ubuf[dst];
        dst ^= 1;
    }
    ctx->launches += nl;
    if (table && !(passes == 1 && table_bits == key_bits)) {
        bucket_table_kernel<<<grid_for(U + 1, 256), 256, 0, st>>>(sorted_key, n_indexed, key_bits - table_bits, table_bits, table);
        LAUNCH_CHECK("bucket_table_kernel");
    }
    return OVL_OK;
}

// ---------------------------------------------------------------- K3
// workspace: [cnt (int64 or 2 x int64) U][scan sums]
size_t ovl_join_workspace_bytes(int64_t n_sources) {
    if (n_sources < 1) n_sources = 1;
    return align256((size_t)n_sources * sizeof(I64x2)) + align256(scan_workspace_bytes(n_sources, sizeof(I64x2)) + scan_workspace_bytes(n_sources, sizeof(int64_t))) + 256;
}

int ovl_join_count(ovl_ctx* ctx, const uint64_t* suffix_key, const uint64_t* prefix_key, const int32_t* len, int32_t k, int64_t U,
                   const uint64_t* sorted_key, const uint32_t* sorted_uid, const int64_t* n_indexed, const int32_t* table,
                   int32_t table_bits, int32_t key_bits, const int32_t* pos_of, const int32_t* copies,
                   const int32_t* sorted_copies, int64_t* cum, int32_t* bucket_lo, int32_t* self_rank, int64_t* pair_off,
                   int64_t* edge_base, void* workspace, size_t workspace_bytes, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !suffix_key || !prefix_key || !len || !sorted_key || !sorted_uid || !n_indexed || !bucket_lo || !self_rank || !pair_off || !workspace)
        return fail(OVL_E_ARG, "ovl_join_count: null argument");
    if (U < 0) return fail(OVL_E_ARG, "ovl_join_count: negative read count");
    if (copies && (!cum || !edge_base)) return fail(OVL_E_ARG, "ovl_join_count: copies given without cum / edge_base");
    if (workspace_bytes < ovl_join_workspace_bytes(U)) return fail(OVL_E_ARG, "ovl_join_count: workspace too small");
    if (key_bits <= 0) key_bits = 2 * k;
    if (table && (table_bits < 1 || table_bits > key_bits)) return fail(OVL_E_ARG, "ovl_join_count: table_bits=%d outside [1, key_bits]", table_bits);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    void* cnt = ws;
    void* sums = ws + align256((size_t)std::max<int64_t>(U, 1) * sizeof(I64x2));
    int nl = 0;
    if (copies) {
        // copies along the sorted index, scanned: the copy mass of any bucket range is a difference of two entries
        if (sorted_copies)
            CUDA_TRY((exclusive_scan<SortedCopiesArray, int64_t>(SortedCopiesArray{sorted_copies}, cum, U, sums, st, &nl)));
        else
            CUDA_TRY((exclusive_scan<SortedCopies, int64_t>(SortedCopies{sorted_uid, copies, n_indexed}, cum, U, sums, st, &nl)));
    }
    if (U > 0) {
        join_count_kernel<<<grid_for(U, 256), 256, 0, st>>>(suffix_key, prefix_key, len, k, U, sorted_key, sorted_uid, n_indexed, table,
                                                             table ? key_bits - table_bits : 0, pos_of, copies, cum, bucket_lo, self_rank,
                                                             copies ? nullptr : (int32_t*)cnt, copies ? (I64x2*)cnt : nullptr);
        LAUNCH_CHECK("join_count_kernel");
    }
    if (copies) {
        CUDA_TRY((exclusive_scan_to<LoadArray<I64x2>, StoreSplit, I64x2>(LoadArray<I64x2>{(const I64x2*)cnt}, StoreSplit{pair_off, edge_base}, U, sums, st, &nl)));
    } else {
        CUDA_TRY((exclusive_scan<LoadArray<int32_t>, int64_t>(LoadArray<int32_t>{(const int32_t*)cnt}, pair_off, U, sums, st, &nl)));
    }
    ctx->launches += nl;
    return OVL_OK;
}

int32_t ovl_totals_len(void) { return kTotalsLen; }

int ovl_join_finalize(ovl_ctx* ctx, const int64_t* pair_off, const int64_t* edge_base, const int32_t* bucket_lo, const int32_t* self_rank,
                      const int64_t* cum, const int32_t* copies, int64_t U, const int32_t* bad_count, const int64_t* n_indexed,
                      int32_t rank, int32_t world, int64_t* totals, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !pair_off || !totals) return fail(OVL_E_ARG, "ovl_join_finalize: null argument");
    if (edge_base && (!bucket_lo || !self_rank || !cum || !copies)) return fail(OVL_E_ARG, "ovl_join_finalize: edge_base given without the join index");
    if (world < 1 || rank < 0 || rank >= world) return fail(OVL_E_ARG, "ovl_join_finalize: rank %d outside world %d", rank, world);
    JoinEdgeIndex jx{pair_off, edge_base, bucket_lo, self_rank, cum, 0, 0};
    join_finalize_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(jx, copies, U, bad_count, n_indexed, rank, world, totals);
    LAUNCH_CHECK("join_finalize_kernel");
    return OVL_OK;
}

int ovl_join_fill(ovl_ctx* ctx, const int64_t* pair_off, int64_t a_begin, int64_t a_end, const int32_t* bucket_lo,
                  const int32_t* self_rank, const uint32_t* sorted_uid, int64_t p_begin, int64_t p_count, int64_t total_hint,
                  int32_t* pair_a, int32_t* pair_b, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !pair_off || !bucket_lo || !self_rank || !sorted_uid || !pair_a || !pair_b) return fail(OVL_E_ARG, "ovl_join_fill: null argument");
    if (p_count <= 0) return OVL_OK;
    int64_t nA = a_end - a_begin;
    // lanes per source read, by the mean bucket size (total_hint / nA); OVL_FILL_LANES overrides (tuning knob)
    cudaStream_t st = (cudaStream_t)stream;
    static const int force_lanes = getenv("OVL_FILL_LANES") ? atoi(getenv("OVL_FILL_LANES")) : -1;
#define FILL_GROUP(L) join_fill_group_kernel<L><<<grid_for(nA * L, 256), 256, 0, st>>>(pair_off, nA, a_begin, bucket_lo, self_rank, sorted_uid, p_begin, p_count, pair_a, pair_b)
    const int lanes = force_lanes >= 0 ? force_lanes
                    : total_hint >= 16 * nA ? 32 : total_hint >= 3 * nA ? 4 : total_hint * 8 >= nA ? 1 : 0;     // measured at 8 M reads: a whole warp per
                                                                                                       // source wins down to ~16 candidates
    if (lanes == 32) FILL_GROUP(32);
    else if (lanes == 8) FILL_GROUP(8);
    else if (lanes == 4) FILL_GROUP(4);
    else if (lanes == 1) FILL_GROUP(1);
    else {
        // sparse: most source reads have no candidate at all -- one thread per output pair
        join_fill_kernel<<<grid_for(p_count, kFillTile), kFillThreads, 0, st>>>(
            pair_off, nA, a_begin, bucket_lo, self_rank, sorted_uid, p_begin, p_count, pair_a, pair_b);
    }
#undef FILL_GROUP
    LAUNCH_CHECK("join_fill kernel");
    return OVL_OK;
}

// ---------------------------------------------------------------- K0-K3 in one call
// Everything between the ASCII reads and the sized pair list, launched back to back from C (the
// per-call cost of going through Python for a dozen 5-50 us kernels was most of the stage's time).
int ovl_candidates_layout(int64_t U, int32_t max_len, int32_t k, int32_t n_segments, int32_t has_copies, ovl_cand_layout* out) {
    if (!out) return fail(OVL_E_ARG, "ovl_candidates_layout: out is null");
    if (U < 0 || max_len < 0) return fail(OVL_E_ARG, "ovl_candidates_layout: negative size");
    if (k < 1 || k > OVL_MAX_K) return fail(OVL_E_UNSUPPORTED, "ovl_candidates_layout: k=%d outside 1..%d", k, OVL_MAX_K);
    if (max_len > OVL_MAX_LONG_READ_LEN) return fail(OVL_E_UNSUPPORTED, "read length %d exceeds the supported maximum %d", max_len, OVL_MAX_LONG_READ_LEN);
    int seg_bits = 0;
    if (n_segments > 1) while (((int64_t)1 << seg_bits) < n_segments) ++seg_bits;
    int key_bits = 2 * k + seg_bits;
    if (key_bits > 64) return fail(OVL_E_UNSUPPORTED, "k=%d with %d read sets needs %d key bits (> 64)", k, n_segments, key_bits);
    memset(out, 0, sizeof(*out));
    int64_t n = std::max<int64_t>(U, 1);
    out->row_words = ovl_row_words(max_len);
    out->key_bits = key_bits;
    out->table_bits = ovl_index_table_bits(U, key_bits);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    out->packed = take((size_t)n * out->row_words * 4 + 16);
    out->len = take((size_t)n * 4);
    out->bad = take(8);
    out->n_indexed = take(8);
    out->prefix_key = take((size_t)n * 8);
    out->suffix_key = take((size_t)n * 8);
    out->sorted_key = take((size_t)n * 8);
    out->sorted_uid = take((size_t)n * 4);
    out->table = take((((size_t)1 << out->table_bits) + 1) * 4);
    out->pos_of = 0;                 // not built by the one-call job: a read's own slot is searched for only when its
                                     // prefix and suffix keys coincide (rare), which saves U scattered 4-byte stores
    out->bucket_lo = take((size_t)n * 4);
    out->self_rank = take((size_t)n * 4);
    out->pair_off = take((size_t)(n + 1) * 8);
    out->edge_base = has_copies ? take((size_t)(n + 1) * 8) : 0;
    out->cum = has_copies ? take((size_t)(n + 1) * 8) : 0;
    out->sorted_copies = has_copies ? take((size_t)n * 4) : 0;
    out->has_copies = has_copies ? 1 : 0;
    out->scratch = off;
    out->scratch_bytes = std::max(ovl_index_workspace_bytes(U), ovl_join_workspace_bytes(U));
    off += align256(out->scratch_bytes);
    out->total_bytes = off + 256;
    return OVL_OK;
}

int ovl_candidates_build(ovl_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int64_t U, int32_t k, const int32_t* segments,
                         const int32_t* copies, int32_t rank, int32_t world, void* arena, const ovl_cand_layout* lay, int64_t* totals,
                         void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !arena || !lay || !totals || (U > 0 && (!ascii || !offsets))) return fail(OVL_E_ARG, "ovl_candidates_build: null argument");
    if ((uintptr_t)arena & 255) return fail(OVL_E_ARG, "ovl_candidates_build: arena must be 256-byte aligned");
    if ((copies != nullptr) != (lay->has_copies != 0)) return fail(OVL_E_ARG, "ovl_candidates_build: copies do not match the layout");
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)arena;
    uint32_t* packed = (uint32_t*)(base + lay->packed);
    int32_t* len = (int32_t*)(base + lay->len);
    int32_t* bad = (int32_t*)(base + lay->bad);
    int64_t* n_indexed = (int64_t*)(base + lay->n_indexed);
    uint64_t* pk = (uint64_t*)(base + lay->prefix_key);
    uint64_t* sk = (uint64_t*)(base + lay->suffix_key);
    uint64_t* skey = (uint64_t*)(base + lay->sorted_key);
    uint32_t* suid = (uint32_t*)(base + lay->sorted_uid);
    int32_t* table = (int32_t*)(base + lay->table);
    int32_t* pos_of = nullptr;
    int32_t* lo = (int32_t*)(base + lay->bucket_lo);
    int32_t* sr = (int32_t*)(base + lay->self_rank);
    int64_t* pair_off = (int64_t*)(base + lay->pair_off);
    int64_t* edge_base = copies ? (int64_t*)(base + lay->edge_base) : nullptr;
    int64_t* cum = copies ? (int64_t*)(base + lay->cum) : nullptr;
    int32_t* sorted_copies = copies ? (int32_t*)(base + lay->sorted_copies) : nullptr;
    void* scratch = base + lay->scratch;
    CUDA_TRY(cudaMemsetAsync(bad, 0, 8, st));
    int rc = pack_launch(ctx, ascii, offsets, U, lay->row_words, k, segments, packed, len, bad, pk, sk, st, "ovl_candidates_build");
    if (rc != OVL_OK) return rc;
    rc = ovl_index_build(ctx, pk, len, U, k, lay->key_bits, skey, suid, n_indexed, table, lay->table_bits, pos_of, copies, sorted_copies,
                         scratch, lay->scratch_bytes, stream);
    if (rc != OVL_OK) return rc;
    rc = ovl_join_count(ctx, sk, pk, len, k, U, skey, suid, n_indexed, table, lay->table_bits, lay->key_bits, pos_of, copies, sorted_copies,
                        cum, lo, sr, pair_off, edge_base, scratch, lay->scratch_bytes, stream);
    if (rc != OVL_OK) return rc;
    return ovl_join_finalize(ctx, pair_off, edge_base, lo, sr, cum, copies, U, bad, n_indexed, rank, world, totals, stream);
}

int ovl_join_count_verify(ovl_ctx* ctx, const uint32_t* packed, int32_t row_words, const int32_t* len, int32_t k,
                          const uint64_t* suffix_hash, int64_t a_begin, int64_t a_end, const uint64_t* sorted_hash,
                          const uint32_t* sorted_uid, const int64_t* n_indexed, int64_t* pair_off, void* workspace,
                          size_t workspace_bytes, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !packed || !len || !suffix_hash || !sorted_hash || !sorted_uid || !n_indexed || !pair_off || !workspace)
        return fail(OVL_E_ARG, "ovl_join_count_verify: null argument");
    int64_t nA = a_end - a_begin;
    if (nA < 0) return fail(OVL_E_ARG, "ovl_join_count_verify: a_end < a_begin");
    if (workspace_bytes < ovl_join_workspace_bytes(nA)) return fail(OVL_E_ARG, "ovl_join_count_verify: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int64_t* cnt = (int64_t*)ws;
    void* sums = ws + align256((size_t)std::max<int64_t>(nA, 1) * sizeof(int64_t));
    if (nA > 0) {
        join_verify_kernel<false><<<grid_for(nA, 256), 256, 0, st>>>(packed, row_words, len, k, suffix_hash, nA, a_begin, sorted_hash,
                                                                     sorted_uid, n_indexed, cnt, nullptr, 0, 0, nullptr, nullptr);
        LAUNCH_CHECK("join_verify_kernel<count>");
    }
    int nl = 0;
    CUDA_TRY((exclusive_scan<LoadArray<int64_t>, int64_t>(LoadArray<int64_t>{cnt}, pair_off, nA, sums, st, &nl)));
    ctx->launches += nl;
    return OVL_OK;
}

int ovl_join_fill_verify(ovl_ctx* ctx, const uint32_t* packed, int32_t row_words, const int32_t* len, int32_t k,
                         const uint64_t* suffix_hash, int64_t a_begin, int64_t a_end, const uint64_t* sorted_hash,
                         const uint32_t* sorted_uid, const int64_t* n_indexed, const int64_t* pair_off, int64_t p_begin,
                         int64_t p_count, int32_t* pair_a, int32_t* pair_b, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !packed || !len || !suffix_hash || !sorted_hash || !sorted_uid || !n_indexed || !pair_off || !pair_a || !pair_b)
        return fail(OVL_E_ARG, "ovl_join_fill_verify: null argument");
    int64_t nA = a_end - a_begin;
    if (nA <= 0 || p_count <= 0) return OVL_OK;
    join_verify_kernel<true><<<grid_for(nA, 256), 256, 0, (cudaStream_t)stream>>>(packed, row_words, len, k, suffix_hash, nA, a_begin,
                                                                                   sorted_hash, sorted_uid, n_indexed, nullptr,
                                                                                   pair_off, p_begin, p_count, pair_a, pair_b);
    LAUNCH_CHECK("join_verify_kernel<fill>");
    return OVL_OK;
}

int ovl_all_pairs_fill(ovl_ctx* ctx, int64_t U, int64_t a_begin, int64_t p_begin, int64_t p_count, int32_t* pair_a,
                       int32_t* pair_b, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !pair_a || !pair_b) return fail(OVL_E_ARG, "ovl_all_pairs_fill: null argument");
    if (p_count <= 0) return OVL_OK;
    if (U < 2) return fail(OVL_E_ARG, "ovl_all_pairs_fill: need at least two reads");
    all_pairs_fill_kernel<<<grid_for(p_count, 256), 256, 0, (cudaStream_t)stream>>>(U, a_begin, p_begin, p_count, pair_a, pair_b);
    LAUNCH_CHECK("all_pairs_fill_kernel");
    return OVL_OK;
}

// ---------------------------------------------------------------- K4 / K5
}  // extern "C"

namespace {

struct DpPlan {
    int mode;       // 1 packed16, 2 int32
    int G, T;
    DpParams prm;
};

#ifndef OVL_DP_EXP_2X76
#define OVL_DP_EXP_2X76 0
#endif
const int kPackedT[] = {19, 25, 32, 38, 76};   // 76 columns per lane only with 16 or 32 lanes (reads > 1,216 bases)
const int kScalarT[] = {32};
const int kLanes[] = {1, 2, 4, 8, 16, 32};

// Range analysis for the cost-space recurrence (see dp.cuh).  N = longest s (rows), M = G*T
// columns.  C[i][j] = beta + i*maxs - H[i][j].  With indel <= 0:
//   max(maxs,0)*min(i,j) >= H[i][j] >= min(base,0)*min(i,j)      (base = min(match, mismatch))
// so 0 <= C <= Cmax = beta + max(maxs,0)*N + |min(base,0)|*min(N,M), beta = max(-maxs,0)*N, and
// the diagonal candidate obeys the same bound.  A gap cost above Cmax can never win and is
// replaced by Cmax+1 (exact).  The packed kernel needs every addend >= 0 and no carry out of a
// 16-bit half: Cmax + max(gu, gl) <= 65535.
bool dp_params(int64_t match, int64_t mismatch, int64_t indel, int64_t N, int64_t M, bool packed, DpParams* out) {
    const int64_t big = 1ll << 40;
    if (std::llabs(match) > big || std::llabs(mismatch) > big) return false;
    int64_t g = std::min<int64_t>(std::max<int64_t>(indel, -big), big);
    int64_t maxs = std::max(match, mismatch), base = std::min(match, mismatch);
    int64_t eqc = maxs - match, nec = maxs - mismatch;
    int64_t mn = std::min(N, M);
    int64_t gu = maxs - g, gl = -g, beta, hi;
    if (g <= 0) {
        beta = std::max<int64_t>(-maxs, 0) * N;
        int64_t cmax = beta + std::max<int64_t>(maxs, 0) * N + std::max<int64_t>(-base, 0) * mn;
        if (gu > cmax) gu = cmax + 1;
        if (gl > cmax) gl = cmax + 1;
        hi = cmax + std::max<int64_t>(std::max(gu, gl), 0);
        if (packed) {
            if (eqc > 127 || nec > 127 || gu < 0 || gl < 0 || hi > 65535 || cmax > 32767) return false;
        } else if (hi >= (1ll << 30) || std::min(gu, gl) <= -(1ll << 30)) {
            return false;
        }
    } else {
        // gaps are rewarded (nothing the assembler uses, but the signature allows it): int32 only
        if (packed) return false;
        beta = 0;
        int64_t mag = (int64_t)std::llabs(maxs) * N + std::max<int64_t>(std::max<int64_t>(std::llabs(match), std::llabs(mismatch)), g) * (N + M);
        if (mag >= (1ll << 30)) return false;
    }
    if (eqc >= (1ll << 30) || nec >= (1ll << 30)) return false;
    const int64_t cmax_out = g <= 0 ? beta + std::max<int64_t>(maxs, 0) * N + std::max<int64_t>(-base, 0) * mn : 0;
    const bool never_out = g <= 0 && (maxs - g) > cmax_out && (-g) > cmax_out;
    out->eqc = (int32_t)eqc; out->nec = (int32_t)nec; out->maxs = (int32_t)maxs; out->beta = (int32_t)beta;
    out->gu = (int32_t)gu; out->gl = (int32_t)gl; out->one = 1u;
    out->cmax = (int32_t)std::min<int64_t>(cmax_out, INT32_MAX); out->gaps_never_win = never_out ? 1 : 0;
    if (packed) {
        out->gu2 = ((uint32_t)gu & 0xffffu) * 0x10001u;
        out->gl2 = ((uint32_t)gl & 0xffffu) * 0x10001u;
        out->maxs2 = (uint32_t)(int32_t)maxs * 0x10001u;
        out->beta2 = ((uint32_t)beta & 0xffffu) * 0x10001u;
    } else {
        out->gu2 = (uint32_t)(int32_t)gu; out->gl2 = (uint32_t)(int32_t)gl;
        out->maxs2 = (uint32_t)(int32_t)maxs; out->beta2 = (uint32_t)(int32_t)beta;
    }
    return true;
}

bool dp_plan(int32_t max_len, int64_t match, int64_t mismatch, int64_t indel, int mode, int forceG, int forceT, DpPlan* plan) {
    int64_t N = std::max(max_len, 1);
    for (int m = 1; m <= 2; ++m) {
        if (mode != 0 && mode != m) continue;
        const int* Ts = m == 1 ? kPackedT : kScalarT;
        int nT = m == 1 ? (int)(sizeof(kPackedT) / sizeof(int)) : 1;
        int64_t best = -1;
        for (int gi = 0; gi < 6; ++gi) {
            for (int ti = 0; ti < nT; ++ti) {
                int G = kLanes[gi], T = Ts[ti];
                if (forceG && G != forceG) continue;
                if (forceT && T != forceT) continue;
#if OVL_DP_EXP_2X76
                if (T == 76 && G < 16 && !(forceG == 2 && forceT == 76)) continue;      // experiment: 2 lanes x 76 columns when forced
#else
                if (T == 76 && G < 16) continue;
#endif
                if ((int64_t)G * T < N) continue;
                DpParams prm;
                if (!dp_params(match, mismatch, indel, N, (int64_t)G * T, m == 1, &prm)) continue;
                int64_t cost = (int64_t)G * T * (N + G - 1);
                if (best < 0 || cost < best) { best = cost; plan->mode = m; plan->G = G; plan->T = T; plan->prm = prm; }
            }
        }
        if (best >= 0) return true;
    }
    return false;
}

template <int G, int T, bool PK, int BITS = 2, bool IMMG = false>
int launch_dp(ovl_ctx* ctx, const uint32_t* packed, int32_t row_words, const int32_t* len, const int32_t* pair_a, const int32_t* pair_b,
              int64_t P, int32_t max_len, const DpParams& prm, int32_t* score, int32_t* end, const DpEdgeOut& eo, cudaStream_t st) {
    constexpr int PAIRS = PK ? 2 : 1;
    constexpr int GROUPS_PER_CTA = (kDpThreads / 32) * (32 / G);
    int64_t groups = (P + PAIRS - 1) / PAIRS;
    int64_t grid = (groups + GROUPS_PER_CTA - 1) / GROUPS_PER_CTA;
    if (grid > 0x7fffffffll) return fail(OVL_E_ARG, "ovl_overlap_dp: too many pairs for one launch (%lld)", (long long)P);
    // the bulk prologue of the packed 2-bit kernels writes table rows [0, G*T) without a bound check
    int lut_rows = dp_lut_rows(dp_bulk<G, T, PK, BITS>() ? std::max(max_len, G * T) : max_len);
    size_t smem = (size_t)GROUPS_PER_CTA * 2 * PAIRS * row_words * sizeof(uint32_t)       // TMA-staged read rows
                + (size_t)GROUPS_PER_CTA * lut_rows * sizeof(uint2)                       // per-row score tables
                + (kDpThreads / 32) * sizeof(uint64_t);                                   // one mbarrier per warp
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(overlap_dp_kernel<G, T, PK, BITS, IMMG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(OVL_E_CUDA, "cudaFuncSetAttribute(smem=%zu) failed: %s", smem, cudaGetErrorString(e));
    }
    overlap_dp_kernel<G, T, PK, BITS, IMMG><<<(unsigned)grid, kDpThreads, smem, st>>>(packed, row_words, len, pair_a, pair_b, P, lut_rows, prm, score, end, eo);
    LAUNCH_CHECK("overlap_dp_kernel");
    return OVL_OK;
}

#define DP_CASE(G_, T_, PK_) \
    if (plan.G == G_ && plan.T == T_) return launch_dp<G_, T_, PK_>(ctx, packed, row_words, len, pair_a, pair_b, P, max_len, plan.prm, score, end, eo, st);
#define DP_CASE_IMM(G_, T_) \
    if (plan.G == G_ && plan.T == T_) return launch_dp<G_, T_, true, 2, true>(ctx, packed, row_words, len, pair_a, pair_b, P, max_len, plan.prm, score, end, eo, st);
#define DP_CASES_IMM_T(T_) DP_CASE_IMM(1, T_) DP_CASE_IMM(2, T_) DP_CASE_IMM(4, T_) DP_CASE_IMM(8, T_) DP_CASE_IMM(16, T_) DP_CASE_IMM(32, T_)
#define DP_CASE8(G_, T_) \
    if (plan.G == G_ && plan.T == T_) return launch_dp<G_, T_, true, 8>(ctx, packed, row_words, len, pair_a, pair_b, P, max_len, plan.prm, score, end, eo, st);
#define DP_CASES8_T(T_) DP_CASE8(1, T_) DP_CASE8(2, T_) DP_CASE8(4, T_) DP_CASE8(8, T_) DP_CASE8(16, T_) DP_CASE8(32, T_)
#define DP_CASES_T(T_, PK_) \
    DP_CASE(1, T_, PK_) DP_CASE(2, T_, PK_) DP_CASE(4, T_, PK_) DP_CASE(8, T_, PK_) DP_CASE(16, T_, PK_) DP_CASE(32, T_, PK_)

}  // namespace

extern "C" {

int ovl_overlap_dp_plan(int32_t max_len, int64_t match, int64_t mismatch, int64_t indel, int32_t mode, int32_t out[3]) {
    if (max_len > OVL_MAX_LONG_READ_LEN)
        return fail(OVL_E_UNSUPPORTED, "overlap DP: read length %d exceeds the supported maximum %d", max_len, OVL_MAX_LONG_READ_LEN);
    if (max_len > OVL_MAX_READ_LEN) {
        DpParams prm;
        if (mode == 1 || !dp_params(match, mismatch, indel, max_len, max_len, false, &prm))
            return fail(OVL_E_UNSUPPORTED, "overlap DP: scores for (match=%lld, mismatch=%lld, indel=%lld, len=%d) do not fit the long-read kernel",
                        (long long)match, (long long)mismatch, (long long)indel, max_len);
        out[0] = 3; out[1] = 0; out[2] = 0;        // mode 3: CTA-per-pair anti-diagonal kernel
        return OVL_OK;
    }
    DpPlan plan;
    if (!dp_plan(max_len, match, mismatch, indel, mode, 0, 0, &plan))
        return fail(OVL_E_UNSUPPORTED, "overlap DP: scores for (match=%lld, mismatch=%lld, indel=%lld, len=%d) do not fit the %s kernels",
                    (long long)match, (long long)mismatch, (long long)indel, max_len, mode == 1 ? "16-bit" : "32-bit");
    out[0] = plan.mode; out[1] = plan.G; out[2] = plan.T;
    return OVL_OK;
}

}  // extern "C"

static int dp_dispatch(ovl_ctx* ctx, const uint32_t* packed, int32_t row_words, const int32_t* len, const int32_t* pair_a,
                       const int32_t* pair_b, int64_t P, int32_t max_len, int64_t match, int64_t mismatch, int64_t indel,
                       int32_t* score, int32_t* end, const DpEdgeOut& eo, int32_t mode, int32_t group_lanes,
                       int32_t cols_per_lane, void* stream, const char* who, int code_bits = 2) {
    if (!ctx || !packed || !len || !pair_a || !pair_b) return fail(OVL_E_ARG, "%s: null argument", who);
    if (P <= 0) return OVL_OK;
    if (row_words < 4 || (row_words & 3) || max_len > (code_bits == 2 ? 16 : 4) * row_words)
        return fail(OVL_E_ARG, "%s: row_words=%d does not hold max_len=%d", who, row_words, max_len);
    if (max_len > OVL_MAX_LONG_READ_LEN)
        return fail(OVL_E_UNSUPPORTED, "%s: read length %d exceeds the supported maximum %d", who, max_len, OVL_MAX_LONG_READ_LEN);
    if (code_bits == 8 && max_len <= 32 * 38 && mode != 2) {
        // byte-coded reads: the packed wavefront kernel with XOR/min/IMAD costs, when the scores fit 16 bits
        DpPlan plan;
        if (dp_plan(max_len, match, mismatch, indel, 1, group_lanes, cols_per_lane, &plan) && plan.T != 19 && plan.T != 76 &&
            plan.prm.eqc == 0) {
            cudaStream_t st = (cudaStream_t)stream;
            DP_CASES8_T(25) DP_CASES8_T(32) DP_CASES8_T(38)
        }
    }
    if (max_len > OVL_MAX_READ_LEN || code_bits == 8) {
        // longer than the register wavefront, or byte-coded reads the packed kernel cannot take:
        // CTA-per-pair anti-diagonal kernel (int32 cost space)
        DpParams prm;
        if (mode == 1 || !dp_params(match, mismatch, indel, max_len, max_len, false, &prm))
            return fail(OVL_E_UNSUPPORTED, "%s: no kernel for (match=%lld, mismatch=%lld, indel=%lld, len=%d, mode=%d)",
                        who, (long long)match, (long long)mismatch, (long long)indel, max_len, mode);
        if (P > 0x7fffffffll) return fail(OVL_E_ARG, "%s: too many pairs for one launch (%lld)", who, (long long)P);
        size_t smem = (size_t)3 * (max_len + 1) * sizeof(int32_t);
        if (code_bits == 8) {
            if (smem > 48 * 1024)
                CUDA_TRY(cudaFuncSetAttribute(overlap_dp_long_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            overlap_dp_long_kernel<8><<<(unsigned)P, kDpLongThreads, smem, (cudaStream_t)stream>>>(packed, row_words, len, pair_a, pair_b, P,
                                                                                                 prm, score, end, eo);
        } else {
            if (smem > 48 * 1024)
                CUDA_TRY(cudaFuncSetAttribute(overlap_dp_long_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            overlap_dp_long_kernel<2><<<(unsigned)P, kDpLongThreads, smem, (cudaStream_t)stream>>>(packed, row_words, len, pair_a, pair_b, P,
                                                                                                 prm, score, end, eo);
        }
        LAUNCH_CHECK("overlap_dp_long_kernel");
        return OVL_OK;
    }
    DpPlan plan;
    if (!dp_plan(max_len, match, mismatch, indel, mode, group_lanes, cols_per_lane, &plan))
        return fail(OVL_E_UNSUPPORTED, "%s: no kernel for (match=%lld, mismatch=%lld, indel=%lld, len=%d, mode=%d, lanes=%d, cols=%d)",
                    who, (long long)match, (long long)mismatch, (long long)indel, max_len, mode, group_lanes, cols_per_lane);
    cudaStream_t st = (cudaStream_t)stream;
    if (plan.mode == 1) {
        // no gap can ever win (the default call site): the instantiation with immediate gap costs
        static const bool imm_ok = !(getenv("OVL_DP_IMM_GAPS") && atoi(getenv("OVL_DP_IMM_GAPS")) == 0);
        if (imm_ok && plan.prm.gaps_never_win && plan.prm.cmax <= (int32_t)kGapNever - 1 && plan.prm.cmax + (int64_t)kGapNever <= 65535) {
            DP_CASES_IMM_T(19) DP_CASES_IMM_T(25) DP_CASES_IMM_T(32) DP_CASES_IMM_T(38)
            DP_CASE_IMM(16, 76) DP_CASE_IMM(32, 76)
#if OVL_DP_EXP_2X76
            DP_CASE_IMM(2, 76)
#endif
        }
        DP_CASES_T(19, true) DP_CASES_T(25, true) DP_CASES_T(32, true) DP_CASES_T(38, true)
        DP_CASE(16, 76, true) DP_CASE(32, 76, true)
    } else {
        DP_CASES_T(32, false)
    }
    return fail(OVL_E_UNSUPPORTED, "%s: instantiation G=%d T=%d mode=%d missing", who, plan.G, plan.T, plan.mode);
}

extern "C" {

int ovl_overlap_dp(ovl_ctx* ctx, const uint32_t* packed, int32_t row_words, const int32_t* len, const int32_t* pair_a,
                   const int32_t* pair_b, int64_t P, int32_t max_len, int64_t match, int64_t mismatch, int64_t indel,
                   int32_t* score, int32_t* end, int32_t mode, int32_t group_lanes, int32_t cols_per_lane, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (P > 0 && (!score || !end)) return fail(OVL_E_ARG, "ovl_overlap_dp: null output");
    DpEdgeOut eo{nullptr, nullptr, nullptr, nullptr, JoinEdgeIndex{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0}};
    return dp_dispatch(ctx, packed, row_words, len, pair_a, pair_b, P, max_len, match, mismatch, indel, score, end, eo,
                       mode, group_lanes, cols_per_lane, stream, "ovl_overlap_dp");
}

int ovl_overlap_dp8(ovl_ctx* ctx, const uint8_t* rows, int32_t row_words, const int32_t* len, const int32_t* pair_a,
                    const int32_t* pair_b, int64_t P, int32_t max_len, int64_t match, int64_t mismatch, int64_t indel,
                    int32_t* score, int32_t* end, const int32_t* copies, const int64_t* node_off, const int64_t* edge_off,
                    int32_t* edges, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (P > 0 && !edges && (!score || !end)) return fail(OVL_E_ARG, "ovl_overlap_dp8: null output");
    if ((uintptr_t)edges & 15) return fail(OVL_E_ARG, "ovl_overlap_dp8: edges must be 16-byte aligned");
    if (copies && (!node_off || !edge_off)) return fail(OVL_E_ARG, "ovl_overlap_dp8: copies given without node_off / edge_off");
    DpEdgeOut eo{(int4*)edges, copies, node_off, edge_off, JoinEdgeIndex{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0}};
    return dp_dispatch(ctx, (const uint32_t*)rows, row_words, len, pair_a, pair_b, P, max_len, match, mismatch, indel, score, end, eo,
                       0, 0, 0, stream, "ovl_overlap_dp8", 8);
}

int ovl_overlap_dp_edges(ovl_ctx* ctx, const uint32_t* packed, int32_t row_words, const int32_t* len, const int32_t* pair_a,
                         const int32_t* pair_b, int64_t P, int32_t max_len, int64_t match, int64_t mismatch, int64_t indel,
                         const int32_t* copies, const int64_t* node_off, const int64_t* edge_off, int32_t* edges,
                         void* stream) {
    ON_CTX_DEVICE(ctx);
    if (P > 0 && !edges) return fail(OVL_E_ARG, "ovl_overlap_dp_edges: null output");
    if ((uintptr_t)edges & 15) return fail(OVL_E_ARG, "ovl_overlap_dp_edges: edges must be 16-byte aligned");
    if (copies && (!node_off || !edge_off)) return fail(OVL_E_ARG, "ovl_overlap_dp_edges: copies given without node_off / edge_off");
    DpEdgeOut eo{(int4*)edges, copies, node_off, edge_off, JoinEdgeIndex{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0}};
    return dp_dispatch(ctx, packed, row_words, len, pair_a, pair_b, P, max_len, match, mismatch, indel, nullptr, nullptr, eo,
                       0, 0, 0, stream, "ovl_overlap_dp_edges");
}

int ovl_overlap_dp_edges_join(ovl_ctx* ctx, const uint32_t* packed, int32_t row_words, const int32_t* len, const int32_t* pair_a,
                              const int32_t* pair_b, int64_t P, int32_t max_len, int64_t match, int64_t mismatch, int64_t indel,
                              const int32_t* copies, const int64_t* node_off, const int64_t* pair_off, const int64_t* edge_base,
                              const int32_t* bucket_lo, const int32_t* self_rank, const int64_t* cum, int64_t p_begin,
                              int64_t e_begin, int32_t* edges, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (P > 0 && !edges) return fail(OVL_E_ARG, "ovl_overlap_dp_edges_join: null output");
    if ((uintptr_t)edges & 15) return fail(OVL_E_ARG, "ovl_overlap_dp_edges_join: edges must be 16-byte aligned");
    if (!copies || !node_off || !pair_off || !edge_base || !bucket_lo || !self_rank || !cum)
        return fail(OVL_E_ARG, "ovl_overlap_dp_edges_join: null join index");
    DpEdgeOut eo{(int4*)edges, copies, node_off, nullptr, JoinEdgeIndex{pair_off, edge_base, bucket_lo, self_rank, cum, p_begin, e_begin}};
    return dp_dispatch(ctx, packed, row_words, len, pair_a, pair_b, P, max_len, match, mismatch, indel, nullptr, nullptr, eo,
                       0, 0, 0, stream, "ovl_overlap_dp_edges_join");
}

// ---------------------------------------------------------------- byte-coded reads (any alphabet)
int ovl_pack_bytes(ovl_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int64_t U, int32_t row_words, uint8_t* rows,
                   int32_t* len, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !ascii || !offsets || !rows || !len) return fail(OVL_E_ARG, "ovl_pack_bytes: null argument");
    if (U <= 0) return OVL_OK;
    if (row_words < 4 || (row_words & 3)) return fail(OVL_E_ARG, "ovl_pack_bytes: row_words must be a positive multiple of 4");
    pack_bytes_kernel<<<grid_for(U * row_words, 256), 256, 0, (cudaStream_t)stream>>>(ascii, offsets, U, row_words * 4, rows, len);
    LAUNCH_CHECK("pack_bytes_kernel");
    return OVL_OK;
}

int ovl_kmer_hashes8(ovl_ctx* ctx, const uint8_t* rows, int32_t row_words, const int32_t* len, int64_t U, int32_t k,
                     uint64_t* prefix_hash, uint64_t* suffix_hash, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !rows || !len || !prefix_hash || !suffix_hash) return fail(OVL_E_ARG, "ovl_kmer_hashes8: null argument");
    if (k < 1) return fail(OVL_E_ARG, "ovl_kmer_hashes8: k must be positive");
    if (U <= 0) return OVL_OK;
    kmer_hash8_kernel<<<grid_for(U, 256), 256, 0, (cudaStream_t)stream>>>(rows, row_words * 4, len, U, k, prefix_hash, suffix_hash);
    LAUNCH_CHECK("kmer_hash8_kernel");
    return OVL_OK;
}

int ovl_join_count_verify8(ovl_ctx* ctx, const uint8_t* rows, int32_t row_words, const int32_t* len, int32_t k,
                           const uint64_t* suffix_hash, int64_t a_begin, int64_t a_end, const uint64_t* sorted_hash,
                           const uint32_t* sorted_uid, const int64_t* n_indexed, int64_t* pair_off, void* workspace,
                           size_t workspace_bytes, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !rows || !len || !suffix_hash || !sorted_hash || !sorted_uid || !n_indexed || !pair_off || !workspace)
        return fail(OVL_E_ARG, "ovl_join_count_verify8: null argument");
    int64_t nA = a_end - a_begin;
    if (nA < 0) return fail(OVL_E_ARG, "ovl_join_count_verify8: a_end < a_begin");
    if (workspace_bytes < ovl_join_workspace_bytes(nA)) return fail(OVL_E_ARG, "ovl_join_count_verify8: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int64_t* cnt = (int64_t*)ws;
    void* sums = ws + align256((size_t)std::max<int64_t>(nA, 1) * sizeof(int64_t));
    if (nA > 0) {
        join_verify8_kernel<false><<<grid_for(nA, 256), 256, 0, st>>>(rows, row_words * 4, len, k, suffix_hash, nA, a_begin, sorted_hash,
                                                                      sorted_uid, n_indexed, cnt, nullptr, 0, 0, nullptr, nullptr);
        LAUNCH_CHECK("join_verify8_kernel<count>");
    }
    int nl = 0;
    CUDA_TRY((exclusive_scan<LoadArray<int64_t>, int64_t>(LoadArray<int64_t>{cnt}, pair_off, nA, sums, st, &nl)));
    ctx->launches += nl;
    return OVL_OK;
}

int ovl_join_fill_verify8(ovl_ctx* ctx, const uint8_t* rows, int32_t row_words, const int32_t* len, int32_t k,
                          const uint64_t* suffix_hash, int64_t a_begin, int64_t a_end, const uint64_t* sorted_hash,
                          const uint32_t* sorted_uid, const int64_t* n_indexed, const int64_t* pair_off, int64_t p_begin,
                          int64_t p_count, int32_t* pair_a, int32_t* pair_b, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !rows || !len || !suffix_hash || !sorted_hash || !sorted_uid || !n_indexed || !pair_off || !pair_a || !pair_b)
        return fail(OVL_E_ARG, "ovl_join_fill_verify8: null argument");
    int64_t nA = a_end - a_begin;
    if (nA <= 0 || p_count <= 0) return OVL_OK;
    join_verify8_kernel<true><<<grid_for(nA, 256), 256, 0, (cudaStream_t)stream>>>(rows, row_words * 4, len, k, suffix_hash, nA, a_begin,
                                                                                    sorted_hash, sorted_uid, n_indexed, nullptr,
                                                                                    pair_off, p_begin, p_count, pair_a, pair_b);
    LAUNCH_CHECK("join_verify8_kernel<fill>");
    return OVL_OK;
}

// ---------------------------------------------------------------- K6
size_t ovl_expand_workspace_bytes(int64_t P) {
    if (P < 1) P = 1;
    return align256(scan_workspace_bytes(P, sizeof(int64_t))) + 256;
}

int ovl_expand_count(ovl_ctx* ctx, const int32_t* pair_a, const int32_t* pair_b, const int32_t* copies, int64_t P,
                     int64_t* edge_off, void* workspace, size_t workspace_bytes, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !edge_off || !workspace || (P > 0 && (!pair_a || !pair_b || !copies))) return fail(OVL_E_ARG, "ovl_expand_count: null argument");
    if (workspace_bytes < ovl_expand_workspace_bytes(P)) return fail(OVL_E_ARG, "ovl_expand_count: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    void* sums = (void*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    // the per-pair edge count copies[a]*copies[b] is computed inside the scan: no count array
    int nl = 0;
    CUDA_TRY((exclusive_scan<CopyProduct, int64_t>(CopyProduct{pair_a, pair_b, copies}, edge_off, P, sums, st, &nl)));
    ctx->launches += nl;
    return OVL_OK;
}

int ovl_expand_fill(ovl_ctx* ctx, const int64_t* edge_off, int64_t P, const int32_t* pair_a, const int32_t* pair_b,
                    const int32_t* score, const int32_t* end, const int32_t* copies, const int64_t* node_off,
                    int64_t e_begin, int64_t e_count, int32_t* edges, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !edge_off || !pair_a || !pair_b || !score || !end || !copies || !node_off || !edges) return fail(OVL_E_ARG, "ovl_expand_fill: null argument");
    if (e_count <= 0) return OVL_OK;
    if ((uintptr_t)edges & 15) return fail(OVL_E_ARG, "ovl_expand_fill: edges must be 16-byte aligned");
    expand_fill_kernel<<<grid_for(e_count, kFillTile), kFillThreads, 0, (cudaStream_t)stream>>>(
        edge_off, P, pair_a, pair_b, score, end, copies, node_off, e_begin, e_count, (int4*)edges);
    LAUNCH_CHECK("expand_fill_kernel");
    return OVL_OK;
}

int ovl_expand_unit(ovl_ctx* ctx, const int32_t* pair_a, const int32_t* pair_b, const int32_t* score, const int32_t* end,
                    int64_t P, int32_t* edges, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !pair_a || !pair_b || !score || !end || !edges) return fail(OVL_E_ARG, "ovl_expand_unit: null argument");
    if (P <= 0) return OVL_OK;
    if ((uintptr_t)edges & 15) return fail(OVL_E_ARG, "ovl_expand_unit: edges must be 16-byte aligned");
    expand_unit_kernel<<<grid_for(P, 256), 256, 0, (cudaStream_t)stream>>>(pair_a, pair_b, score, end, P, (int4*)edges);
    LAUNCH_CHECK("expand_unit_kernel");
    return OVL_OK;
}

// ---------------------------------------------------------------- edge filter (all-pairs builders)
size_t ovl_filter_workspace_bytes(int64_t E) {
    if (E < 1) E = 1;
    return align256(scan_workspace_bytes(E, sizeof(int64_t))) + 256;
}

int ovl_filter_count(ovl_ctx* ctx, const int32_t* edges, int64_t E, int32_t min_weight, int64_t* keep_off, void* workspace,
                     size_t workspace_bytes, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !keep_off || !workspace || (E > 0 && !edges)) return fail(OVL_E_ARG, "ovl_filter_count: null argument");
    if (workspace_bytes < ovl_filter_workspace_bytes(E)) return fail(OVL_E_ARG, "ovl_filter_count: workspace too small");
    void* sums = (void*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int nl = 0;
    CUDA_TRY((exclusive_scan<EdgeKept, int64_t>(EdgeKept{(const int4*)edges, min_weight}, keep_off, E, sums, (cudaStream_t)stream, &nl)));
    ctx->launches += nl;
    return OVL_OK;
}

int ovl_filter_fill(ovl_ctx* ctx, const int32_t* edges, const int64_t* keep_off, int64_t E, int32_t min_weight, int32_t* out,
                    void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || (E > 0 && (!edges || !keep_off || !out))) return fail(OVL_E_ARG, "ovl_filter_fill: null argument");
    if (E <= 0) return OVL_OK;
    filter_edges_kernel<<<grid_for(E, 256), 256, 0, (cudaStream_t)stream>>>((const int4*)edges, keep_off, E, min_weight, (int4*)out);
    LAUNCH_CHECK("filter_edges_kernel");
    return OVL_OK;
}

// ---------------------------------------------------------------- K7
size_t ovl_align_pair_workspace_bytes(int32_t n, int32_t m) {
    if (n < 0) n = 0;
    if (m < 0) m = 0;
    size_t ints = (size_t)3 * (n + 1) + (size_t)(m + 1);
    return align256(ints * sizeof(int32_t)) + align256((size_t)(n + 1) * (m + 1)) + 256;
}

int ovl_align_pair(ovl_ctx* ctx, const int32_t* s, int32_t n, const int32_t* t, int32_t m, int64_t match, int64_t mismatch,
                   int64_t indel, void* workspace, size_t workspace_bytes, int32_t* result, uint8_t* ops, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !workspace || !result || !ops || (n > 0 && !s) || (m > 0 && !t)) return fail(OVL_E_ARG, "ovl_align_pair: null argument");
    if (n < 0 || m < 0) return fail(OVL_E_ARG, "ovl_align_pair: negative length");
    if (workspace_bytes < ovl_align_pair_workspace_bytes(n, m)) return fail(OVL_E_ARG, "ovl_align_pair: workspace too small");
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int32_t* diag = (int32_t*)ws;
    int32_t* last_row = diag + (size_t)3 * (n + 1);
    int8_t* tb = (int8_t*)(ws + align256(((size_t)3 * (n + 1) + (size_t)(m + 1)) * sizeof(int32_t)));
    int threads = std::min(kAlignThreads, std::max(32, ((std::max(std::min(n, m), 1) + 31) / 32) * 32));
    size_t smem = (size_t)3 * (n + 1) * sizeof(int32_t);
    if (smem <= 200 * 1024) {
        if (smem > 48 * 1024)
            CUDA_TRY(cudaFuncSetAttribute(align_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        align_pair_kernel<true><<<1, threads, smem, (cudaStream_t)stream>>>(s, n, t, m, match, mismatch, indel, diag, last_row, tb, result, ops);
    } else {
        align_pair_kernel<false><<<1, threads, 0, (cudaStream_t)stream>>>(s, n, t, m, match, mismatch, indel, diag, last_row, tb, result, ops);
    }
    LAUNCH_CHECK("align_pair_kernel");
    return OVL_OK;
}

// ---------------------------------------------------------------- K8
size_t ovl_local_align_workspace_bytes(int32_t n, int32_t m) {
    if (n < 0) n = 0;
    if (m < 0) m = 0;
    // the row-per-thread kernel stores its traceback diagonal-major: (n + m + 1) diagonals of n + 1 bytes
    return align256((size_t)3 * (n + 1) * sizeof(int32_t)) + align256((size_t)(n + 1) * (n + m + 2)) + 256;
}

int ovl_local_align(ovl_ctx* ctx, const int32_t* query, int32_t n, const int32_t* reference, int32_t m, int64_t match,
                    int64_t mismatch, int64_t indel, void* workspace, size_t workspace_bytes, int32_t* result, uint8_t* ops,
                    void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !workspace || !result || !ops || (n > 0 && !query) || (m > 0 && !reference)) return fail(OVL_E_ARG, "ovl_local_align: null argument");
    if (n < 0 || m < 0) return fail(OVL_E_ARG, "ovl_local_align: negative length");
    if (workspace_bytes < ovl_local_align_workspace_bytes(n, m)) return fail(OVL_E_ARG, "ovl_local_align: workspace too small");
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int32_t* diag = (int32_t*)ws;
    int8_t* tb = (int8_t*)(ws + align256((size_t)3 * (n + 1) * sizeof(int32_t)));
    if (n >= 1 && n <= kAlignThreads) {
        // row-per-thread sweep: neighbours exchange cells with shuffles, one barrier per diagonal
        local_align_rows_kernel<<<1, ((n + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(query, n, reference, m, match, mismatch, indel,
                                                                                    tb, result, ops);
        LAUNCH_CHECK("local_align_rows_kernel");
        return OVL_OK;
    }
    // as many threads as the longest anti-diagonal needs (a block barrier per diagonal: fewer warps, cheaper barrier)
    int threads = std::min(kAlignThreads, std::max(32, ((std::min(n, m) + 31) / 32) * 32));
    size_t smem = (size_t)3 * (n + 1) * sizeof(int32_t);
    if (smem <= 200 * 1024) {
        if (smem > 48 * 1024)
            CUDA_TRY(cudaFuncSetAttribute(local_align_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        local_align_kernel<true><<<1, threads, smem, (cudaStream_t)stream>>>(query, n, reference, m, match, mismatch, indel, diag, tb, result, ops);
    } else {
        local_align_kernel<false><<<1, threads, 0, (cudaStream_t)stream>>>(query, n, reference, m, match, mismatch, indel, diag, tb, result, ops);
    }
    LAUNCH_CHECK("local_align_kernel");
    return OVL_OK;
}

int ovl_local_align_batch(ovl_ctx* ctx, const int32_t* queries, const int64_t* q_off, int32_t n_queries, int32_t max_query_len,
                          const int32_t* reference, const int32_t* ref_start, const int32_t* ref_len, int64_t match,
                          int64_t mismatch, int64_t indel, uint8_t* tb, const int64_t* tb_off, int32_t* results, uint8_t* ops,
                          const int64_t* ops_off, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (n_queries <= 0) return OVL_OK;
    if (!ctx || !queries || !q_off || !reference || !ref_start || !ref_len || !tb || !tb_off || !results || !ops || !ops_off)
        return fail(OVL_E_ARG, "ovl_local_align_batch: null argument");
    if (max_query_len < 0 || max_query_len > kAlignThreads)
        return fail(OVL_E_UNSUPPORTED, "ovl_local_align_batch: queries longer than %d symbols go through ovl_local_align", kAlignThreads);
    int threads = std::max(32, ((max_query_len + 31) / 32) * 32);
    local_align_batch_kernel<<<(unsigned)n_queries, threads, 0, (cudaStream_t)stream>>>(
        queries, q_off, reference, ref_start, ref_len, match, mismatch, indel, (int8_t*)tb, tb_off, results, ops, ops_off);
    LAUNCH_CHECK("local_align_batch_kernel");
    return OVL_OK;
}

// ---------------------------------------------------------------- read simulator
size_t ovl_simulate_workspace_bytes(int64_t n_reads) {
    if (n_reads < 1) n_reads = 1;
    return align256((size_t)n_reads * 8) * 2 + align256(scan_workspace_bytes(n_reads, sizeof(int64_t))) + 256;
}

int ovl_simulate_reads(ovl_ctx* ctx, const uint8_t* genome, int64_t genome_len, int64_t n_reads, int32_t read_len, uint32_t error_thr,
                       uint64_t seed, int64_t* offsets, uint8_t* ascii, void* workspace, size_t workspace_bytes, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !genome || !offsets || !ascii || !workspace) return fail(OVL_E_ARG, "ovl_simulate_reads: null argument");
    if (genome_len < 1 || genome_len >= (1ll << 32) || read_len < 1 || n_reads < 0)
        return fail(OVL_E_ARG, "ovl_simulate_reads: need 1 <= genome_len < 2^32, read_len >= 1, n_reads >= 0");
    if (workspace_bytes < ovl_simulate_workspace_bytes(n_reads)) return fail(OVL_E_ARG, "ovl_simulate_reads: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int64_t* start = (int64_t*)ws;                 ws += align256((size_t)std::max<int64_t>(n_reads, 1) * 8);
    int64_t* len = (int64_t*)ws;                   ws += align256((size_t)std::max<int64_t>(n_reads, 1) * 8);
    void* sums = ws;
    int nl = 0;
    if (n_reads > 0) {
        sim_starts_kernel<<<grid_for(n_reads, 256), 256, 0, st>>>(n_reads, genome_len, read_len, seed, start, len);
        LAUNCH_CHECK("sim_starts_kernel");
    }
    CUDA_TRY((exclusive_scan<LoadArray<int64_t>, int64_t>(LoadArray<int64_t>{len}, offsets, n_reads, sums, st, &nl)));
    ctx->launches += nl;
    if (n_reads > 0) {
        int64_t chunks = (read_len + 15) / 16;
        sim_bases_kernel<<<grid_for(n_reads * chunks, 256), 256, 0, st>>>(genome, n_reads, read_len, error_thr, seed, start, offsets, ascii);
        LAUNCH_CHECK("sim_bases_kernel");
    }
    return OVL_OK;
}

// ---------------------------------------------------------------- cycle-removal pre-pass
size_t ovl_trim_workspace_bytes(int64_t n_nodes) {
    if (n_nodes < 1) n_nodes = 1;
    return align256((size_t)n_nodes * 4) + 512;
}

int ovl_trim_sinks(ovl_ctx* ctx, const int32_t* src, const int32_t* dst, int64_t E, int64_t n_nodes, int32_t* state, void* workspace,
                   size_t workspace_bytes, int32_t* h_rounds, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !state || !workspace || (E > 0 && (!src || !dst))) return fail(OVL_E_ARG, "ovl_trim_sinks: null argument");
    if (n_nodes < 0 || E < 0) return fail(OVL_E_ARG, "ovl_trim_sinks: negative size");
    if (workspace_bytes < ovl_trim_workspace_bytes(n_nodes)) return fail(OVL_E_ARG, "ovl_trim_sinks: workspace too small");
    if (h_rounds) *h_rounds = 0;
    if (n_nodes == 0) return OVL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    int32_t* last_change = (int32_t*)ws;
    int32_t* outdeg = (int32_t*)(ws + 256);
    CUDA_TRY(cudaMemsetAsync(ws, 0, 256 + (size_t)n_nodes * 4, st));
    if (E > 0) {
        trim_outdeg_kernel<<<grid_for(E, 256), 256, 0, st>>>(src, E, outdeg);
        LAUNCH_CHECK("trim_outdeg_kernel");
    }
    trim_init_kernel<<<grid_for(n_nodes, 256), 256, 0, st>>>(n_nodes, outdeg, state);
    LAUNCH_CHECK("trim_init_kernel");
    // peel in batches of rounds; stop when a whole batch passed without any node dying
    int32_t round = 1;
    const int kBatch = 32;
    while (E > 0) {
        for (int i = 0; i < kBatch; ++i, ++round) {
            trim_round_kernel<<<grid_for(E, 256), 256, 0, st>>>(src, dst, E, outdeg, state, round, last_change);
            LAUNCH_CHECK("trim_round_kernel");
        }
        int32_t last = 0;
        CUDA_TRY(cudaMemcpyAsync(&last, last_change, sizeof(last), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (last < round) break;                        // nobody died in the batch's last round: the peeling has stopped
        if (round > n_nodes + kBatch) return fail(OVL_E_CUDA, "ovl_trim_sinks: peeling did not converge");
    }
    if (h_rounds) *h_rounds = round - 1;
    return OVL_OK;
}

// ---------------------------------------------------------------- edge-list fingerprint
int ovl_edge_list_hash(ovl_ctx* ctx, const int32_t* edges, int64_t E, int64_t first_row, uint64_t* accum, void* stream) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !accum || (E > 0 && !edges)) return fail(OVL_E_ARG, "ovl_edge_list_hash: null argument");
    if (E <= 0) return OVL_OK;
    if ((uintptr_t)edges & 15) return fail(OVL_E_ARG, "ovl_edge_list_hash: edges must be 16-byte aligned");
    unsigned grid = (unsigned)std::min<int64_t>((E + 255) / 256, (int64_t)ctx->sm_count * 16);
    edge_hash_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const int4*)edges, E, first_row, (unsigned long long*)accum);
    LAUNCH_CHECK("edge_hash_kernel");
    return OVL_OK;
}

// ---------------------------------------------------------------- probe
}  // extern "C"

template <int KIND>
static int run_probe(ovl_ctx* ctx, int iters, double* gops, double* ms_out) {
    if (!ctx->probe_sink) CUDA_TRY(cudaMalloc(&ctx->probe_sink, 256));
    int blocks = ctx->sm_count * 8;     // 2048 threads per SM: every scheduler has 16 warps
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    int_probe_kernel<KIND><<<blocks, kProbeThreads>>>(ctx->probe_sink, 4, 0x00070003u, 0xfff5fff5u, 1u);   // warm-up
    CUDA_TRY(cudaEventRecord(e0));
    int_probe_kernel<KIND><<<blocks, kProbeThreads>>>(ctx->probe_sink, iters, 0x00070003u, 0xfff5fff5u, 1u);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    LAUNCH_CHECK("int_probe_kernel");
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double ops = (double)blocks * kProbeThreads * (double)iters * kProbeUnroll * kProbeChains * probe_ops_per_iter(KIND);
    *gops = ops / (ms * 1e-3) / 1e9;
    if (ms_out) *ms_out = ms;
    return OVL_OK;
}

extern "C" {

int ovl_int_peak_probe(ovl_ctx* ctx, int32_t kind, int32_t iters, double* h_gops, double* h_ms) {
    ON_CTX_DEVICE(ctx);
    if (!ctx || !h_gops) return fail(OVL_E_ARG, "ovl_int_peak_probe: null argument");
    if (iters < 1) iters = 1;
    switch (kind) {
        case 0: return run_probe<0>(ctx, iters, h_gops, h_ms);
        case 1: return run_probe<1>(ctx, iters, h_gops, h_ms);
        case 2: return run_probe<2>(ctx, iters, h_gops, h_ms);
        case 3: return run_probe<3>(ctx, iters, h_gops, h_ms);
        case 4: return run_probe<4>(ctx, iters, h_gops, h_ms);
        case 5: return run_probe<5>(ctx, iters, h_gops, h_ms);
        case 6: return run_probe<6>(ctx, iters, h_gops, h_ms);
        case 7: return run_probe<7>(ctx, iters, h_gops, h_ms);
        case 8: return run_probe<8>(ctx, iters, h_gops, h_ms);
        case 9: return run_probe<9>(ctx, iters, h_gops, h_ms);
        case 10: return run_probe<10>(ctx, iters, h_gops, h_ms);
        case 11: return run_probe<11>(ctx, iters, h_gops, h_ms);
        case 12: return run_probe<12>(ctx, iters, h_gops, h_ms);
        case 13: return run_probe<13>(ctx, iters, h_gops, h_ms);
        case 14: return run_probe<14>(ctx, iters, h_gops, h_ms);
        case 15: return run_probe<15>(ctx, iters, h_gops, h_ms);
        case 16: return run_probe<16>(ctx, iters, h_gops, h_ms);
        case 17: return run_probe<17>(ctx, iters, h_gops, h_ms);
        case 18: return run_probe<18>(ctx, iters, h_gops, h_ms);
        case 20: return run_probe<20>(ctx, iters, h_gops, h_ms);
        case 21: return run_probe<21>(ctx, iters, h_gops, h_ms);
        case 22: return run_probe<22>(ctx, iters, h_gops, h_ms);
        case 23: return run_probe<23>(ctx, iters, h_gops, h_ms);
        case 24: return run_probe<24>(ctx, iters, h_gops, h_ms);
        case 25: return run_probe<25>(ctx, iters, h_gops, h_ms);
        case 26: return run_probe<26>(ctx, iters, h_gops, h_ms);
        case 27: return run_probe<27>(ctx, iters, h_gops, h_ms);
        case 28: return run_probe<28>(ctx, iters, h_gops, h_ms);
        case 29: return run_probe<29>(ctx, iters, h_gops, h_ms);
        case 30: return run_probe<30>(ctx, iters, h_gops, h_ms);
        default: return fail(OVL_E_ARG, "ovl_int_peak_probe: unknown kind %d", kind);
    }
}

}  // extern "C"
