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
// The (pair, rule end) items are flattened over the lanes so that the dependent look-ups
// (rule end -> node -> row bitmap) run 32 wide; only items whose row is non-zero touch the arena.
template <typename CT, typename F>
__device__ __forceinline__ void scan_entity(const rl_graph &g, const rl_rules &r, const CT *__restrict__ arena, size_t abase,
                                            const uint32_t *__restrict__ mbase, int hc0, const int32_t *__restrict__ tp, int e, F f)
{
    const int lane = threadIdx.x & 31;
    const int p0 = g.ent_ptr[e], p1 = g.ent_ptr[e + 1];
    for (int pb = p0; pb < p1; pb += 32) {
        const int pi = pb + lane;
        int row = 0, t0 = 0, cnt = 0;
        if (pi < p1) {
            const int rel = g.ent_rel[pi];
            row = g.ent_row[pi];
            t0 = tp[rel];
            cnt = tp[rel + 1] - t0;
        }
        int P = cnt;                                        // inclusive scan: items of pairs 0..lane
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, P, o);
            if (lane >= o) P += t;
        }
        const int T = __shfl_sync(FULL, P, 31);
        const int first = P - cnt;
        for (int base = 0; base < T; base += 32) {
            const int k = min(base + lane, T - 1);
            int pr = 0;                                      // pair of item k: #lanes with P <= k
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (__shfl_sync(FULL, P, pr + step - 1) <= k) pr += step;
            const int t = __shfl_sync(FULL, t0, pr) + (k - __shfl_sync(FULL, first, pr));
            const int rw = __shfl_sync(FULL, row, pr);
            long long addr = -1;
            if (base + lane < T) {
                const int v = __ldg(r.term_node + t);
                if (row_valid(r.node_chunk0, mbase, hc0, v, rw)) addr = (long long)r.node_row_off[v] + rw;
            }
            uint32_t live = __ballot_sync(FULL, addr >= 0);
            while (live) {
                const int j = __ffs(live) - 1;
                live &= live - 1;
                const long long a = __shfl_sync(FULL, addr, j);
                const int tj = __shfl_sync(FULL, t, j);
                f(arena[(abase + (size_t)a) * RL_LANES + lane], tj);
            }
        }
    }
}
