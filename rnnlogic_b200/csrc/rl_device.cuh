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
#define SCORE_WARPS_MAX 8         // k_predictor_scores: warps (= entity words) per block, one softmax partial per block
#ifndef SCORE_MIN_BLOCKS
#define SCORE_MIN_BLOCKS 4      // 64 registers, no spills: swept 3..8 x 2..8 rows on the B200
#endif
#ifndef SCORE_ROWS
#define SCORE_ROWS 4            // count rows a k_predictor_scores warp keeps in flight (8 spills at 64 registers)
#endif
#ifndef BWD_ROWS
#define BWD_ROWS 8               // (count row, G row) pairs a k_predictor_bwd_stream warp keeps in flight
#endif
#define CE_WARPS 32             // k_ce_finalize / k_grad_sparse: one warp per query lane of the slot

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
__device__ __forceinline__ float warp_maxf(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
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

// per group (= one reference batch, possibly several slots): loss = -sum lp*tgt / max(sum tgt, 1)
// (trainer.py:89); slot_invT[s] = 1 / max(T_group, 1)
static __global__ void k_group_reduce(int n_groups, const int32_t *__restrict__ group_ptr, const float *__restrict__ slot_lsum,
                               const float *__restrict__ slot_tsum, float *__restrict__ group_loss,
                               float *__restrict__ group_tsum, float *__restrict__ slot_invT, float *__restrict__ stats)
{
    const int gi = blockIdx.x, lane = threadIdx.x;               // one warp per group, lane = query lane
    if (gi >= n_groups) return;
    const int s0 = group_ptr ? group_ptr[gi] : gi, s1 = group_ptr ? group_ptr[gi + 1] : gi + 1;
    double L = 0.0, T = 0.0;
    for (int k = s0; k < s1; ++k) { L += (double)slot_lsum[k]; T += (double)slot_tsum[k]; }
    const float Tf = fmaxf((float)T, 1.f);
    if (lane == 0) {
        group_tsum[gi] = (float)T;
        group_loss[gi] = (float)L / Tf;
    }
    for (int k = s0; k < s1; ++k) {
        if (lane == 0) slot_invT[k] = 1.f / Tf;
        float *st = stats + ((size_t)k * 32 + lane) * 4;          // stats[.][3]: valid flag -> softmax-gradient coefficient
        st[3] = st[3] != 0.f ? st[2] / st[1] / Tf : 0.f;          // S_b / sum-exp / T'
    }
}

// ---- item list helpers -------------------------------------------------------------------------
// k_numeric appends one item {row (slot-relative), t0, entity, n} per NON-ZERO row of every rule-end node
// (the rules ending there are node_term_rule[t0 .. t0 + n)) and counts it in bucket_cnt[slot][entity]; rl_sort_items (k_items_scan + k_items_scatter) turns the
// counts into offsets and groups a slot's items by entity, so the items of entity e are the contiguous
// range [bucket_off[e], bucket_off[e+1]) of items_sorted.  Both tables have RL_BUCKET_STRIDE ints per slot.
#define RL_BUCKET_STRIDE(W) ((size_t)(W) * 32 + 32)

struct WordItems {
    const int4 *its;      // the slot's items, grouped by entity
    const uint32_t *masks; // lane mask of every item (same order), or nullptr
    uint32_t wmask;       // window: lane mask of item wbase + lane (all lanes set when masks == nullptr)
    int b0, b1;           // item range of the lane's entity (entity = word * 32 + lane)
    int wbase, wend;      // window: lane holds item wbase + lane (the word's items are one contiguous range ending at wend)
    int4 win;
    uint32_t present;     // entities of the word that own at least one item
};

__device__ __forceinline__ void word_items_window(WordItems &wi, int at)
{
    const int lane = threadIdx.x & 31;
    wi.wbase = at;
    const bool in = at + lane < wi.wend;
    wi.win = in ? __ldg(wi.its + at + lane) : make_int4(0, 0, -1, 0);
    wi.wmask = wi.masks ? (in ? __ldg(wi.masks + at + lane) : 0u) : FULL;
}

__device__ __forceinline__ WordItems load_word_items(const rl_frontier &fr, const rl_slots &s, int W, int slot, int ew)
{
    const int lane = threadIdx.x & 31;
    WordItems wi;
    const int *off = fr.bucket_off + (size_t)slot * RL_BUCKET_STRIDE(W) + (size_t)ew * 32;
    wi.b0 = off[lane];                                       // padding entities of the last word: empty ranges
    wi.b1 = off[lane + 1];
    wi.its = reinterpret_cast<const int4 *>(fr.items_sorted) + fr.item_off[slot];
    wi.masks = fr.item_mask_sorted ? fr.item_mask_sorted + fr.item_off[slot] : nullptr;
    wi.present = __ballot_sync(FULL, wi.b1 > wi.b0);
    wi.wend = __shfl_sync(FULL, wi.b1, 31);
    word_items_window(wi, __shfl_sync(FULL, wi.b0, 0));      // most words fit one window: one item load per word
    return wi;
}

// f(count_of_this_lane, t) for every (item of entity i of the word, rule t ending at the item's node);
// t indexes rl_rules::node_term_rule.  Called by all 32 lanes (lane = query), four row loads in flight.
// Entities are normally visited in ascending order, so the 32-item window only moves forward.
#ifndef ITEM_ROWS_IN_FLIGHT
#define ITEM_ROWS_IN_FLIGHT 4
#endif
template <typename CT, typename F, int U = ITEM_ROWS_IN_FLIGHT>
__device__ __forceinline__ void for_entity_items(WordItems &wi, const rl_rules &r, const CT *__restrict__ arena_slot,
                                                 int i, F f)
{
    const int lane = threadIdx.x & 31;
    const int i0 = __shfl_sync(FULL, wi.b0, i), i1 = __shfl_sync(FULL, wi.b1, i);
    int j = i0;
    while (j < i1) {
        if (j < wi.wbase || j >= wi.wbase + 32) word_items_window(wi, j);
        const int cnt = min(i1, wi.wbase + 32) - j;          // items of this entity inside the window, from j on
        for (int j0 = 0; j0 < cnt; j0 += U) {
            CT cv[U];
            int tv[U], nv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int src = (j - wi.wbase + j0 + u) & 31;
                const int a = __shfl_sync(FULL, wi.win.x, src);
                const int t0 = __shfl_sync(FULL, wi.win.y, src);
                const int nt = __shfl_sync(FULL, wi.win.w, src);
                tv[u] = t0;
                nv[u] = j0 + u < cnt ? nt : 0;
                cv[u] = j0 + u < cnt ? arena_slot[(size_t)a * RL_LANES + lane] : (CT)0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                for (int t = tv[u]; t < tv[u] + nv[u]; ++t) f(cv[u], t);
        }
        j += cnt;
    }
}
