// Pre-pass for the weakest-edge cycle removal (overlapGraphs.py:106-130), SURVEY 8f-4.
// The reference repeats nx.find_cycle + remove-weakest-edge until the graph is a DAG.  A node that
// cannot REACH any cycle can never lie on the cycle a depth-first search reports, and a search that
// wanders into such a node only comes back empty-handed; dropping those nodes therefore leaves the
// sequence of cycles found -- and of edges removed -- unchanged.  They are exactly the nodes peeled
// off by repeatedly deleting sinks (out-degree 0), computed here on the device edge list:
//   state[v] = r > 0  : v became a sink in round r (cannot reach a cycle)
//   state[v] = 0      : v survives (it reaches a cycle or lies on one)
// Every edge is handled once, in the round after its head died.
#pragma once
#include "common.cuh"

namespace ovl {

__global__ void __launch_bounds__(256) trim_outdeg_kernel(const int32_t* __restrict__ src, int64_t E, int32_t* __restrict__ outdeg) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) atomicAdd(&outdeg[src[e]], 1);
}

__global__ void __launch_bounds__(256) trim_init_kernel(int64_t n, const int32_t* __restrict__ outdeg, int32_t* __restrict__ state) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) state[v] = outdeg[v] == 0 ? 1 : 0;
}

__global__ void __launch_bounds__(256) trim_round_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t E,
                                                         int32_t* __restrict__ outdeg, int32_t* __restrict__ state, int32_t round,
                                                         int32_t* __restrict__ last_change) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    if (state[dst[e]] != round) return;                 // the head died in the previous round: this edge goes now
    int32_t u = src[e];
    if (atomicSub(&outdeg[u], 1) == 1) {                // u has just lost its last out-edge
        state[u] = round + 1;
        atomicMax(last_change, round + 1);
    }
}

}  // namespace ovl
