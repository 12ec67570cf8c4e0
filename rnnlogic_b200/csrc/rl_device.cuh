// Shared device helpers of the rnnlogic_b200 kernels (rl_kernels.cu, rl_plus.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "rnnlogic_b200.h"

#define FULL 0xffffffffu
#define WARPS_PER_BLOCK 8
#define SM_ROWS_PER_BLOCK 512   // entity rows per block in the softmax / rank sweeps

// error state + launch counter live in rl_kernels.cu
int rl_fail(int code, const char *what, cudaError_t e = cudaSuccess);
void rl_count_launch();

#define CHECK_LAUNCH(name)                                             \
    do {                                                               \
        cudaError_t e_ = cudaGetLastError();                           \
        if (e_ != cudaSuccess) return rl_fail(RL_ERR_CUDA, name, e_);  \
        rl_count_launch();                                             \
    } while (0)

__device__ __forceinline__ int rank_row(const rl_graph &g, int rel, int e)
{
    const uint2 w = __ldg(reinterpret_cast<const uint2 *>(g.rank_tab) + (size_t)rel * g.rank_words + (e >> 5));
    const uint32_t bit = 1u << (e & 31);
    return (w.x & bit) ? (int)(w.y + __popc(w.x & (bit - 1))) : -1;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_sumf(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_sumi(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// lower_bound over sorted keys; returns index or -1
__device__ __forceinline__ int find_key(const rl_answers &a, long long key)
{
    long long lo = 0, hi = a.num_keys;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (a.keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < a.num_keys && a.keys[lo] == key) ? (int)lo : -1;
}

__device__ __forceinline__ bool row_valid(const int32_t *__restrict__ node_chunk0, const uint32_t *mbase, int hc0, int v, int row)
{
    return (mbase[node_chunk0[v] - hc0 + (row >> 5)] >> (row & 31)) & 1u;
}

// Warp-cooperative walk over every (non-zero terminal row, rule end) that can contribute to entity e
// of one slot: f(count_of_this_lane, term_index) is called by all 32 lanes (lane = query).  Order is
// fixed: the entity's (relation,row) pairs ascending, then the head's rule ends of that relation.
template <typename CT, typename F>
__device__ __forceinline__ void scan_entity(const rl_graph &g, const rl_rules &r, const CT *__restrict__ arena, size_t abase,
                                            const uint32_t *__restrict__ mbase, int hc0, const int32_t *__restrict__ tp, int e, F f)
{
    const int lane = threadIdx.x & 31;
    const int p0 = g.ent_ptr[e], p1 = g.ent_ptr[e + 1];
    for (int pb = p0; pb < p1; pb += 32) {
        const int pi = pb + lane;
        int row = 0, t0 = 0, t1 = 0;
        if (pi < p1) {
            const int rel = g.ent_rel[pi];
            row = g.ent_row[pi];
            t0 = tp[rel];
            t1 = tp[rel + 1];
        }
        uint32_t have = __ballot_sync(FULL, t1 > t0);
        while (have) {
            const int k = __ffs(have) - 1;
            have &= have - 1;
            const int a0 = __shfl_sync(FULL, t0, k), a1 = __shfl_sync(FULL, t1, k);
            const int rw = __shfl_sync(FULL, row, k);
            for (int t = a0; t < a1; ++t) {
                const int v = __ldg(r.term_node + t);
                if (!row_valid(r.node_chunk0, mbase, hc0, v, rw)) continue;
                f(arena[(abase + (size_t)r.node_row_off[v] + rw) * RL_LANES + lane], t);
            }
        }
    }
}
