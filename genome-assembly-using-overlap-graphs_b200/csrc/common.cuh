// Shared device helpers for the overlap-detection kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// -DOVL_BOUNDS_CHECKS=1 compiles index checks into the kernels (every shared-memory table access and scattered global
// store the index of which is computed from data): a violated check traps, which fails the launch and with it the
// test that made it.  compute-sanitizer is closed on the GPU pool, so the GPU suite is also run once against a library
// built this way (tools/gpu_round23.sh, profiles/r3k_pytest_gpu_bounds_checked.log).
#ifndef OVL_BOUNDS_CHECKS
#define OVL_BOUNDS_CHECKS 0
#endif
#if OVL_BOUNDS_CHECKS
#define OVL_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define OVL_CHECK(cond) do { } while (0)
#endif

namespace ovl {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// streaming 128-bit accesses that do not pollute L1
__device__ __forceinline__ uint4 ldg_nc_v4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_na_v4(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// first index in [lo, hi) with a[idx] >= key
template <typename T>
__device__ __forceinline__ int64_t lower_bound(const T* __restrict__ a, int64_t lo, int64_t hi, T key) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first index in [lo, hi) with a[idx] > key
template <typename T>
__device__ __forceinline__ int64_t upper_bound(const T* __restrict__ a, int64_t lo, int64_t hi, T key) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace ovl
