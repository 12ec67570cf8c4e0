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

// ---- item list helpers -------------------------------------------------------------------------
// k_numeric appends one item {row (slot-relative), node, entity, -} per NON-ZERO row of every rule-end
// node; k_items_sort buckets a slot's items by entity word.  A warp that owns the 32 entities of a word
// loads its bucket once (first 32 items stay in registers) and walks the items of one entity at a time.
struct WordItems {
    const int4 *its;      // the word's bucket inside items_sorted
    int n;                // items in the bucket
    int4 it0;             // lane's item of the first 32 (z = -1: none)
    uint32_t present;     // entities of the word that own at least one item
};

__device__ __forceinline__ WordItems load_word_items(const rl_frontier &fr, const rl_slots &s, int W, int slot, int ew)
{
    const int lane = threadIdx.x & 31;
    WordItems wi;
    const int *off = fr.bucket_off + (size_t)slot * (W + 1);
    const int b0 = off[ew];
    wi.n = off[ew + 1] - b0;
    wi.its = reinterpret_cast<const int4 *>(fr.items_sorted) + fr.item_off[slot] + b0;
    wi.present = 0u;
    wi.it0 = make_int4(0, 0, -1, 0);
    for (int c0 = 0; c0 < wi.n; c0 += 32) {
        const int4 it = c0 + lane < wi.n ? wi.its[c0 + lane] : make_int4(0, 0, -1, 0);
        if (c0 == 0) wi.it0 = it;
        wi.present |= __reduce_or_sync(FULL, it.z >= 0 ? 1u << (it.z & 31) : 0u);
    }
    return wi;
}

// f(count_of_this_lane, t) for every (item of entity i of the word, rule t ending at the item's node);
// t indexes rl_rules::node_term_rule.  Called by all 32 lanes (lane = query), four row loads in flight.
template <typename CT, typename F>
__device__ __forceinline__ void for_entity_items(const WordItems &wi, const rl_rules &r, const CT *__restrict__ arena_slot,
                                                 int i, F f)
{
    const int lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < wi.n; c0 += 32) {
        const int4 it = c0 == 0 ? wi.it0 : (c0 + lane < wi.n ? wi.its[c0 + lane] : make_int4(0, 0, -1, 0));
        uint32_t sel = __ballot_sync(FULL, it.z >= 0 && (it.z & 31) == i);
        while (sel) {
            CT cv[4];
            int nv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = sel ? __ffs(sel) - 1 : -1;
                sel &= sel - 1;
                const int a = __shfl_sync(FULL, it.x, j & 31);
                nv[u] = j >= 0 ? __shfl_sync(FULL, it.y, j & 31) : -1;
                cv[u] = j >= 0 ? arena_slot[(size_t)a * RL_LANES + lane] : (CT)0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (nv[u] < 0) continue;
                for (int t = r.node_term_ptr[nv[u]]; t < r.node_term_ptr[nv[u] + 1]; ++t) f(cv[u], t);
            }
        }
    }
}
