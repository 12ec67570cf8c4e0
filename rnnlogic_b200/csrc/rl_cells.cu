// rnnlogic_b200 -- the scoring half of the hot path on CANDIDATE CELLS instead of dense [N][32] matrices.
//
// A cell is a (query, entity) pair whose total path count is non-zero (src/predictors.py:224-225,239).
// With an entity bias every other logit of a query is just bias[e] (predictors.py:257-262: zeros scattered
// with the candidate scores, plus bias), so the all-entity softmax cross-entropy (src/trainer.py:88-89)
// splits exactly into one per-step scalar pair over all entities
//        Mg = max_e bias[e],   Sg = sum_e exp(bias[e] - Mg)
// plus corrections over a query's cells,
//        S_b = (Sg - sum_cells exp(bias[e] - Mg)) * exp(Mg - M_b) + sum_cells exp(bias[e] + z - M_b),
// and the dense part of the bias gradient is rank one: exp(bias[e] - Mg) * K.  Without an entity feature
// (predictors.py:267-269: -inf outside the mask) the softmax runs over the cells alone.  Nothing of size
// [S][N][32] is written or read: per slot there are two int32 tables of N entries (candidate word + first
// cell) and per cell a score, a gradient and a key.  On the FB15k-237-shape workload 3.3 % of the
// (query, entity) pairs are cells.
//
// Pipeline of a train step (all launches on one stream, no host sync):
//   rl_expand_level (items + lane masks) -> rl_cells_build (group items by entity, OR the lane masks into
//   the candidate words, number the cells) -> scores per cell (rl_predictor_cell_scores, or the
//   PredictorPlus aggregate + MLP kernels) -> rl_cells_softmax_ce (loss + gradient per cell + bias
//   gradient) -> backward (rl_predictor_cell_backward / PredictorPlus kernels).
#include "rl_device.cuh"

#define CELL_BLOCKS 4          // blocks per slot in the per-query sweeps over the cells (softmax partials, rank)
#define PC_WARPS 4             // k_pred_cells / k_pred_cells_bwd: entity words per block
#define PC_ROWS 4              // count rows a warp keeps in flight

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ------------------------------------------------------------------------------------------
// cell numbering: one block per slot.  Cells are numbered slot by slot (base from ONE atomicAdd per
// slot -- the order between slots is arbitrary and internal), entity-major, lane-minor.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
k_cell_scan(rl_graph g, rl_rules r, rl_slots s, rl_frontier fr, rl_cells c)
{
    __shared__ int wsum[16];
    __shared__ int s_base;
    const int slot = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = g.num_entities;
    uint32_t *nz = c.nzmask + (size_t)slot * N;
    int32_t *off = c.cand_off + (size_t)slot * N;
    const int q = s.slot_head[slot];
    if (tid < 32 && r.zr_ptr[q + 1] > r.zr_ptr[q]) {          // empty-body rules: count = one_hot(h) -> (h_b, b) is a cell
        const int h = s.lane_h[slot * RL_LANES + tid];
        if (h >= 0) atomicOr(nz + h, 1u << tid);
    }
    __syncthreads();
    int tot = 0;
    for (int e = tid; e < N; e += 512) tot += __popc(__ldcg(nz + e));
    tot = warp_sumi(tot);
    if (lane == 0) wsum[warp] = tot;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int k = 0; k < 16; ++k) t += wsum[k];
        s_base = atomicAdd(c.counters, t);
        c.slot_ncell[slot] = t;
        if ((long long)s_base + t > (long long)c.cap) c.counters[1] = 1;     // the step must be redone with larger arrays
    }
    __syncthreads();
    int carry = s_base;
    for (int e0 = 0; e0 < N; e0 += 512) {
        const int e = e0 + tid;
        uint32_t bits = e < N ? __ldcg(nz + e) : 0u;
        const int v = __popc(bits);
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int t = wsum[k];
            if (k < warp) wbase += t;
            total += t;
        }
        int idx = carry + wbase + incl - v;
        if (e < N) off[e] = idx;
        while (bits) {                                          // cell -> (slot, lane) keys for the per-cell kernels
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            if (idx < c.cap) c.cell_key[idx] = slot * RL_LANES + b;
            ++idx;
        }
        carry += total;
        __syncthreads();
    }
    if (c.nnz_cap <= 0) return;
    // first non-zero of every entity-grouped item: prefix of the items' lane-mask populations
    const int n = fr.item_cnt[slot];
    const long long ibase = fr.item_off[slot];
    const uint32_t *im = fr.item_mask_sorted + ibase;
    tot = 0;
    for (int i = tid; i < n; i += 512) tot += __popc(im[i]);
    tot = warp_sumi(tot);
    if (lane == 0) wsum[warp] = tot;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int k = 0; k < 16; ++k) t += wsum[k];
        s_base = atomicAdd(c.counters + 2, t);
        if ((long long)s_base + t > (long long)c.nnz_cap) c.counters[3] = 1;
    }
    __syncthreads();
    carry = s_base;
    for (int i0 = 0; i0 < n; i0 += 512) {
        const int i = i0 + tid;
        const int v = i < n ? __popc(im[i]) : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int t = wsum[k];
            if (k < warp) wbase += t;
            total += t;
        }
        if (i < n) c.nnz_off[ibase + i] = carry + wbase + incl - v;
        carry += total;
        __syncthreads();
    }
}

// Mg = max bias, Sg = sum exp(bias - Mg)
__global__ void __launch_bounds__(1024)
k_bias_stats(int N, const float *__restrict__ bias, double *__restrict__ acc)
{
    __shared__ float red_m[32];
    __shared__ double red_s[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float m = -INFINITY;
    for (int e = tid; e < N; e += 1024) m = fmaxf(m, bias[e]);
    m = warp_maxf(m);
    if (lane == 0) red_m[warp] = m;
    __syncthreads();
    m = red_m[0];
    for (int k = 1; k < 32; ++k) m = fmaxf(m, red_m[k]);
    double sum = 0.0;
    for (int e = tid; e < N; e += 1024) sum += (double)expf(bias[e] - m);
    sum = warp_sum(sum);
    if (lane == 0) red_s[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int k = 0; k < 32; ++k) t += red_s[k];
        acc[0] = (double)m;
        acc[1] = t;
        acc[2] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------
// Predictor scores per cell: zc[cell] = sum_rule w_rule * fp32(count) (predictors.py:58-65), without the bias.
// One warp per 32 entities streams the word's items (entity-grouped), PC_ROWS count rows in flight; only the
// lanes of an item's mask touch its count row (32-byte sectors instead of the whole 128-byte line).
// ------------------------------------------------------------------------------------------
template <typename CT>
__global__ void __launch_bounds__(PC_WARPS * 32)
k_pred_cells(rl_graph g, rl_rules r, rl_slots s, rl_frontier fr, rl_cells c, const float *__restrict__ w,
             float *__restrict__ zc)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.y;
    const int ew = blockIdx.x * PC_WARPS + warp;
    const int N = g.num_entities, W = g.rank_words;
    if (ew >= W) return;
    const int e_lane = ew * 32 + lane;
    const uint32_t my_bits = e_lane < N ? c.nzmask[(size_t)slot * N + e_lane] : 0u;
    if (__ballot_sync(FULL, my_bits != 0u) == 0u) return;
    const int my_off = e_lane < N ? c.cand_off[(size_t)slot * N + e_lane] : 0;
    const int q = s.slot_head[slot];
    const CT *arena = reinterpret_cast<const CT *>(fr.arena) + (size_t)s.arena_off[slot] * RL_LANES;
    const int h = s.lane_h[slot * RL_LANES + lane];
    const int z0 = r.zr_ptr[q], z1 = r.zr_ptr[q + 1];
    double zsum = 0.0;
    for (int t = z0; t < z1; ++t) zsum += (double)__ldg(w + r.zr_rule[t]);
    const uint32_t lt = lanemask_lt();
    auto finish = [&](int i, double acc) {
        if (z1 > z0 && h == ew * 32 + i) acc += zsum;         // empty-body rules: count = one_hot(h)
        const uint32_t bits = __shfl_sync(FULL, my_bits, i);
        const int off = __shfl_sync(FULL, my_off, i);
        if ((bits >> lane) & 1u) {
            const int idx = off + __popc(bits & lt);
            if (idx < c.cap) zc[idx] = (float)acc;
        }
    };
    WordItems wi = load_word_items(fr, s, W, slot, ew);
    int cur = -1;
    double acc = 0.0;
    const bool coo = c.nnz_cap > 0;
    const long long ibase = fr.item_off[slot];
    const int B0 = __shfl_sync(FULL, wi.b0, 0), B1 = wi.wend;
    for (int c0 = B0; c0 < B1; c0 += 32) {
        if (c0 != wi.wbase) word_items_window(wi, c0);
        const int cnt = min(32, B1 - c0);
        const int my_n0 = (coo && lane < cnt) ? c.nnz_off[ibase + c0 + lane] : 0;
        for (int j0 = 0; j0 < cnt; j0 += PC_ROWS) {
            CT cv[PC_ROWS];
            float wv[PC_ROWS];
#pragma unroll
            for (int u = 0; u < PC_ROWS; ++u) {
                const int src = (j0 + u) & 31;
                const int a = __shfl_sync(FULL, wi.win.x, src);
                const int t0 = __shfl_sync(FULL, wi.win.y, src);
                const uint32_t m = __shfl_sync(FULL, wi.wmask, src);
                const bool ok = j0 + u < cnt;
                cv[u] = (ok && ((m >> lane) & 1u)) ? arena[(size_t)a * RL_LANES + lane] : (CT)0;
                wv[u] = ok ? __ldg(w + __ldg(r.node_term_rule + t0)) : 0.f;
                if (coo) {                                           // coordinate list for the backward: (cell, item, fp32 count)
                    const int n0 = __shfl_sync(FULL, my_n0, src);
                    const int ie = __shfl_sync(FULL, wi.win.z, src) & 31;
                    const uint32_t bits = __shfl_sync(FULL, my_bits, ie);
                    const int off = __shfl_sync(FULL, my_off, ie);
                    if (ok && ((m >> lane) & 1u)) {
                        const int pos = n0 + __popc(m & lt);
                        if (pos < c.nnz_cap) {
                            c.nz_val[pos] = (float)cv[u];
                            c.nz_cell[pos] = off + __popc(bits & lt);
                            c.nz_item[pos] = (int)(ibase + c0 + j0 + u);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < PC_ROWS; ++u) {
                if (j0 + u >= cnt) break;
                const int src = (j0 + u) & 31;
                const int i = __shfl_sync(FULL, wi.win.z, src) & 31;
                const int nt = __shfl_sync(FULL, wi.win.w, src);
                if (i != cur) {
                    if (cur >= 0) finish(cur, acc);
                    cur = i;
                    acc = 0.0;
                }
                const double cf = (double)(float)cv[u];                                // x.float() * w (predictors.py:64)
                acc += cf * (double)wv[u];
                if (nt > 1) {                                                          // duplicate rules ending at the same node
                    const int t0 = __shfl_sync(FULL, wi.win.y, src);
                    for (int t = t0 + 1; t < t0 + nt; ++t) acc += cf * (double)__ldg(w + r.node_term_rule[t]);
                }
            }
        }
    }
    if (cur >= 0) finish(cur, acc);
    const uint32_t present = __ballot_sync(FULL, my_bits != 0u);
    for (uint32_t todo = present & ~wi.present; todo; todo &= todo - 1) finish(__ffs(todo) - 1, 0.0);   // cells of empty-body rules only
}

// logit of (query b of the slot, entity e): bias[e] + cell score, or "no logit" (mask mode, not a cell)
__device__ __forceinline__ bool cell_logit(const rl_cells &c, const float *__restrict__ bias, const float *__restrict__ zc,
                                           size_t srow, int e, int b, float &l, int &idx)
{
    const uint32_t bits = c.nzmask[srow + e];
    idx = -1;
    l = bias ? bias[e] : 0.f;
    if ((bits >> b) & 1u) {
        idx = c.cand_off[srow + e] + __popc(bits & ((1u << b) - 1u));
        if (idx < c.cap) l += zc[idx]; else idx = -1;
        return true;
    }
    return bias != nullptr;
}

// order-preserving float <-> unsigned key (atomicMax on floats of either sign)
__device__ __forceinline__ unsigned fkey(float v)
{
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// One block per slot.  Threads first own ENTITIES (each walks the cells of its entities: 32 entities of a warp in
// flight instead of one): per-query max of the cell logits, then per-query sum of exp(l - M) - exp(bias - M) (the
// correction of the all-entity sum Sg), both through a few replicated shared-memory atomics.  Then one warp per
// query lane walks the sparse smoothed target of its query (data.py:207-212, trainer.py:84) -> loss sums.
// stats[slot][lane] = (M, S, S_b, valid)
#define CE_COPIES 8
__global__ void __launch_bounds__(CE_WARPS * 32)
k_ce_cells(rl_graph g, rl_slots s, rl_cells c, rl_answers ans, float smoothing, const float *__restrict__ bias,
           const double *__restrict__ acc, const float *__restrict__ zc, float *__restrict__ stats,
           float *__restrict__ slot_lsum, float *__restrict__ slot_tsum)
{
    __shared__ unsigned sm_max[CE_COPIES][32];
    __shared__ float sm_sum[CE_COPIES][32];
    __shared__ float sm_M[32], sm_S[32];
    __shared__ double red_l[CE_WARPS], red_t[CE_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, b = tid >> 5;
    const int slot = blockIdx.x;
    const int N = g.num_entities;
    const int q = s.slot_head[slot];
    const size_t srow = (size_t)slot * N;
    const float Mg = bias ? (float)acc[0] : 0.f;
    if (tid < CE_COPIES * 32) {
        (&sm_max[0][0])[tid] = fkey(-INFINITY);
        (&sm_sum[0][0])[tid] = 0.f;
    }
    __syncthreads();
    const int cp = b & (CE_COPIES - 1);
    for (int e = tid; e < N; e += CE_WARPS * 32) {              // pass 1: max
        uint32_t bits = c.nzmask[srow + e];
        if (!bits) continue;
        int idx = c.cand_off[srow + e];
        const float bl = bias ? bias[e] : 0.f;
        for (; bits; bits &= bits - 1, ++idx) {
            const float l = bl + (idx < c.cap ? zc[idx] : 0.f);
            atomicMax(&sm_max[cp][__ffs(bits) - 1], fkey(l));
        }
    }
    __syncthreads();
    if (tid < 32) {
        unsigned k = sm_max[0][tid];
#pragma unroll
        for (int i = 1; i < CE_COPIES; ++i) k = max(k, sm_max[i][tid]);
        const float Mc = fkey_inv(k);
        sm_M[tid] = bias ? fmaxf(Mg, Mc) : Mc;
    }
    __syncthreads();
    for (int e = tid; e < N; e += CE_WARPS * 32) {              // pass 2: sum of the corrections
        uint32_t bits = c.nzmask[srow + e];
        if (!bits) continue;
        int idx = c.cand_off[srow + e];
        const float bl = bias ? bias[e] : 0.f;
        for (; bits; bits &= bits - 1, ++idx) {
            const int qb = __ffs(bits) - 1;
            const float M = sm_M[qb];
            const float l = bl + (idx < c.cap ? zc[idx] : 0.f);
            const float v = expf(l - M) - (bias ? expf(bl - M) : 0.f);
            if (v != 0.f) atomicAdd(&sm_sum[cp][qb], v);
        }
    }
    __syncthreads();
    if (tid < 32) {
        double Sd = 0.0;
#pragma unroll
        for (int i = 0; i < CE_COPIES; ++i) Sd += (double)sm_sum[i][tid];
        if (bias) Sd += acc[1] * (double)expf(Mg - sm_M[tid]);
        sm_S[tid] = (float)Sd;
    }
    __syncthreads();
    const float M = sm_M[b], S = sm_S[b];
    const int h = s.lane_h[slot * RL_LANES + b];
    const int t = s.lane_t[slot * RL_LANES + b];
    float lsum = 0.f, tacc = 0.f, sb = 0.f;
    bool saw_t = false;
    const bool valid = h >= 0 && M != -INFINITY;
    if (valid) {
        auto term = [&](int e, float tg) {
            float l;
            int idx;
            if (!cell_logit(c, bias, zc, srow, e, b, l, idx)) return;       // outside the mask (trainer.py:89)
            const float p = expf(l - M) / S;
            lsum += logf(p + 1e-8f) * tg;
            tacc += tg;
            sb += tg / (p + 1e-8f) * p;
        };
        const int ki = find_key(ans, (long long)q * N + h);
        const int a0 = ki >= 0 ? ans.ptr[ki] : 0, a1 = ki >= 0 ? ans.ptr[ki + 1] : 0;
        for (int a = a0 + lane; a < a1; a += 32) {
            const int e = ans.ent[a];
            float tg = smoothing;
            if (e == t) { tg += 1.f - smoothing; saw_t = true; }
            term(e, tg);
        }
        saw_t = __any_sync(FULL, saw_t);
        if (!saw_t && t >= 0 && lane == 0) term(t, 1.f - smoothing);
    }
    lsum = warp_sumf(lsum);
    tacc = warp_sumf(tacc);
    sb = warp_sumf(sb);
    if (lane == 0) {
        float *st = stats + ((size_t)slot * 32 + b) * 4;
        st[0] = M; st[1] = S; st[2] = sb; st[3] = valid ? 1.f : 0.f;
        red_l[b] = (double)lsum;
        red_t[b] = (double)tacc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double L = 0.0, T = 0.0;
        for (int k = 0; k < CE_WARPS; ++k) { L += red_l[k]; T += red_t[k]; }
        slot_tsum[slot] = (float)T;
        slot_lsum[slot] = (float)(-L);
    }
}

// Gc[cell] = scale * softmax * S_b / T'  (dense part of dloss/dlogit at the cells), and the bias gradient's
// correction at the cells: what the cell's logit contributes beyond the rank-one term exp(bias - M_b) * coef.
// After k_group_reduce: stats[.][3] = S_b / sum-exp / T'.  Threads own entities (one atomic per entity with cells).
__global__ void __launch_bounds__(256)
k_grad_cells(rl_graph g, rl_cells c, const float *__restrict__ bias, double *__restrict__ acc,
             const float *__restrict__ zc, const float *__restrict__ stats, float scale, float *__restrict__ Gc,
             float *__restrict__ grad_bias)
{
    __shared__ float sM[32], sC[32];
    const int tid = threadIdx.x;
    const int slot = blockIdx.y;
    const int N = g.num_entities;
    const size_t srow = (size_t)slot * N;
    if (tid < 32) {
        const float4 st = __ldg(reinterpret_cast<const float4 *>(stats) + (size_t)slot * 32 + tid);
        const float coef = st.w * scale;
        sM[tid] = st.x;
        sC[tid] = coef;
        if (bias && blockIdx.x == 0) {                          // K = sum over queries of coef * exp(Mg - M_b)
            double v = coef != 0.f ? (double)coef * (double)expf((float)acc[0] - st.x) : 0.0;
            v = warp_sum(v);
            if (tid == 0 && v != 0.0) atomicAdd(acc + 2, v);
        }
    }
    __syncthreads();
    for (int e = blockIdx.x * 256 + tid; e < N; e += gridDim.x * 256) {
        uint32_t bits = c.nzmask[srow + e];
        if (!bits) continue;
        int idx = c.cand_off[srow + e];
        const float bl = bias ? bias[e] : 0.f;
        float corr = 0.f;
        for (; bits; bits &= bits - 1, ++idx) {
            if (idx >= c.cap) break;
            const int qb = __ffs(bits) - 1;
            const float coef = sC[qb], M = sM[qb];
            const float gq = coef != 0.f ? expf(bl + zc[idx] - M) * coef : 0.f;
            Gc[idx] = gq;
            if (bias && coef != 0.f) corr += gq - expf(bl - M) * coef;
        }
        if (bias && corr != 0.f) atomicAdd(grad_bias + e, corr);
    }
}

// target terms: dlogit -= scale * p * tgt / (p + eps) / T' at the target entries (one warp per query lane)
__global__ void __launch_bounds__(CE_WARPS * 32)
k_grad_targets(rl_graph g, rl_slots s, rl_cells c, rl_answers ans, float smoothing, const float *__restrict__ bias,
               const float *__restrict__ zc, const float *__restrict__ stats, const float *__restrict__ slot_invT,
               float scale, float *__restrict__ Gc, float *__restrict__ grad_bias)
{
    const int lane = threadIdx.x & 31, b = threadIdx.x >> 5;
    const int slot = blockIdx.x;
    const int N = g.num_entities;
    const int q = s.slot_head[slot];
    const float *st = stats + ((size_t)slot * 32 + b) * 4;
    if (st[3] == 0.f) return;
    const float iT = slot_invT[slot] * scale;
    const int h = s.lane_h[slot * RL_LANES + b];
    const int t = s.lane_t[slot * RL_LANES + b];
    const float M = st[0], S = st[1];
    const size_t srow = (size_t)slot * N;
    auto apply = [&](int e, float tg) {
        float l;
        int idx;
        if (!cell_logit(c, bias, zc, srow, e, b, l, idx)) return;
        const float p = expf(l - M) / S;
        const float term = p * (tg / (p + 1e-8f)) * iT;
        if (idx >= 0) Gc[idx] -= term;
        if (grad_bias) atomicAdd(grad_bias + e, -term);
    };
    const int ki = find_key(ans, (long long)q * N + h);
    const int a0 = ki >= 0 ? ans.ptr[ki] : 0, a1 = ki >= 0 ? ans.ptr[ki + 1] : 0;
    bool saw_t = false;
    for (int a = a0 + lane; a < a1; a += 32) {
        const int e = ans.ent[a];
        float tg = smoothing;
        if (e == t) { tg += 1.f - smoothing; saw_t = true; }
        apply(e, tg);
    }
    saw_t = __any_sync(FULL, saw_t);
    if (!saw_t && t >= 0 && lane == 0) apply(t, 1.f - smoothing);
}

// the rank-one part of the bias gradient: every (query, entity) pair contributes exp(bias[e] - M_b) * coef_b
__global__ void __launch_bounds__(256)
k_bias_finish(int N, const float *__restrict__ bias, const double *__restrict__ acc, float *__restrict__ grad_bias)
{
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= N) return;
    grad_bias[e] += (float)(acc[2] * (double)expf(bias[e] - (float)acc[0]));
}

// ------------------------------------------------------------------------------------------
// Predictor backward from the cells: grad_w[rule] += <Gc[cells of e], fp32(count row)> for every item
// ------------------------------------------------------------------------------------------
template <typename CT>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 4)
k_pred_cells_bwd(rl_graph g, rl_rules r, rl_slots s, rl_frontier fr, rl_cells c, const float *__restrict__ Gc,
                 float *__restrict__ grad_w)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.y;
    const int ew = blockIdx.x * WARPS_PER_BLOCK + warp;
    const int N = g.num_entities, W = g.rank_words;
    const int q = s.slot_head[slot];
    const size_t srow = (size_t)slot * N;
    const int z0 = r.zr_ptr[q], z1 = r.zr_ptr[q + 1];
    if (blockIdx.x == 0 && warp == 0 && z1 > z0) {                 // empty-body rules: count = one_hot(h)
        const int h = s.lane_h[slot * RL_LANES + lane];
        double v = 0.0;
        if (h >= 0) {
            const uint32_t bits = c.nzmask[srow + h];
            const int idx = c.cand_off[srow + h] + __popc(bits & ((1u << lane) - 1u));
            if (((bits >> lane) & 1u) && idx < c.cap) v = (double)Gc[idx];
        }
        v = warp_sum(v);
        if (lane == 0 && v != 0.0)
            for (int t = z0; t < z1; ++t) atomicAdd(grad_w + r.zr_rule[t], (float)v);
    }
    if (ew >= W) return;
    const int e_lane = ew * 32 + lane;
    const uint32_t my_bits = e_lane < N ? c.nzmask[srow + e_lane] : 0u;
    if (__ballot_sync(FULL, my_bits != 0u) == 0u) return;
    const int my_off = e_lane < N ? c.cand_off[srow + e_lane] : 0;
    const CT *arena = reinterpret_cast<const CT *>(fr.arena) + (size_t)s.arena_off[slot] * RL_LANES;
    const uint32_t lt = lanemask_lt();
    WordItems wi = load_word_items(fr, s, W, slot, ew);
    const int B0 = __shfl_sync(FULL, wi.b0, 0), B1 = wi.wend;
    for (int c0 = B0; c0 < B1; c0 += 32) {
        if (c0 != wi.wbase) word_items_window(wi, c0);
        const int cnt = min(32, B1 - c0);
        for (int j0 = 0; j0 < cnt; j0 += BWD_ROWS) {
            float pv[BWD_ROWS];
#pragma unroll
            for (int u = 0; u < BWD_ROWS; ++u) {
                const int src = (j0 + u) & 31;
                const int a = __shfl_sync(FULL, wi.win.x, src);
                const int i = __shfl_sync(FULL, wi.win.z, src) & 31;
                const uint32_t m = __shfl_sync(FULL, wi.wmask, src);
                const uint32_t bits = __shfl_sync(FULL, my_bits, i);
                const int off = __shfl_sync(FULL, my_off, i);
                pv[u] = 0.f;
                if (j0 + u < cnt && ((m >> lane) & 1u)) {          // a non-zero count => the cell exists
                    const int idx = off + __popc(bits & lt);
                    const CT cv = arena[(size_t)a * RL_LANES + lane];
                    if (idx < c.cap) pv[u] = (float)cv * Gc[idx];
                }
            }
#pragma unroll
            for (int u = 0; u < BWD_ROWS; ++u) {
                if (j0 + u >= cnt) break;
                const float v = warp_sumf(pv[u]);
                const int src = (j0 + u) & 31;
                const int t0 = __shfl_sync(FULL, wi.win.y, src), nt = __shfl_sync(FULL, wi.win.w, src);
                if (lane == 0 && v != 0.f)
                    for (int t = t0; t < t0 + nt; ++t) atomicAdd(grad_w + r.node_term_rule[t], v);
            }
        }
    }
}

// The same backward from the coordinate list: one thread per non-zero count, no count row is read again.
__global__ void __launch_bounds__(256)
k_pred_nnz_bwd(rl_rules r, rl_frontier fr, rl_cells c, const float *__restrict__ Gc, float *__restrict__ grad_w)
{
    const int n = min(c.counters[2], c.nnz_cap);
    const int4 *items = reinterpret_cast<const int4 *>(fr.items_sorted);
    for (int p = blockIdx.x * 256 + threadIdx.x; p < n; p += gridDim.x * 256) {
        const int cell = c.nz_cell[p];
        if (cell >= c.cap) continue;
        const float v = c.nz_val[p] * Gc[cell];
        if (v == 0.f) continue;
        const int4 it = __ldg(items + c.nz_item[p]);              // {row, first rule end, entity, rule ends}
        for (int t = it.y; t < it.y + it.w; ++t) atomicAdd(grad_w + __ldg(r.node_term_rule + t), v);
    }
}

// empty-body rules: count = one_hot(h) -> grad_w[rule] += sum over the slot's queries of Gc[cell(h_b, b)]
__global__ void __launch_bounds__(32)
k_pred_zr_bwd(rl_graph g, rl_rules r, rl_slots s, rl_cells c, const float *__restrict__ Gc, float *__restrict__ grad_w)
{
    const int lane = threadIdx.x, slot = blockIdx.x;
    const int q = s.slot_head[slot];
    const int z0 = r.zr_ptr[q], z1 = r.zr_ptr[q + 1];
    if (z1 <= z0) return;
    const size_t srow = (size_t)slot * g.num_entities;
    const int h = s.lane_h[slot * RL_LANES + lane];
    double v = 0.0;
    if (h >= 0) {
        const uint32_t bits = c.nzmask[srow + h];
        const int idx = c.cand_off[srow + h] + __popc(bits & ((1u << lane) - 1u));
        if (((bits >> lane) & 1u) && idx < c.cap) v = (double)Gc[idx];
    }
    v = warp_sum(v);
    if (lane == 0 && v != 0.0)
        for (int t = z0; t < z1; ++t) atomicAdd(grad_w + r.zr_rule[t], (float)v);
}

// ------------------------------------------------------------------------------------------
// filtered rank on the cells (trainer.py:189-201).  With a bias: #{e : logit > val} = #{e : bias[e] > val}
// (binary search in the sorted bias table) + corrections over the query's cells - the known answers.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_rank_cells(rl_graph g, rl_slots s, rl_cells c, const float *__restrict__ bias, const float *__restrict__ zc,
             int32_t *__restrict__ counters)
{
    __shared__ int sm_gt[WARPS_PER_BLOCK][32], sm_ge[WARPS_PER_BLOCK][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.y;
    const int N = g.num_entities, W = g.rank_words;
    const size_t srow = (size_t)slot * N;
    const int t = s.lane_t[slot * RL_LANES + lane];
    float val = INFINITY;
    if (t >= 0) {
        int idx;
        if (!cell_logit(c, bias, zc, srow, t, lane, val, idx)) val = INFINITY;
    }
    const uint32_t lt = lanemask_lt();
    int gt = 0, ge = 0;
    for (int ew = blockIdx.x * WARPS_PER_BLOCK + warp; ew < W; ew += CELL_BLOCKS * WARPS_PER_BLOCK) {
        const int e_lane = ew * 32 + lane;
        const uint32_t my_bits = e_lane < N ? c.nzmask[srow + e_lane] : 0u;
        uint32_t present = __ballot_sync(FULL, my_bits != 0u);
        if (!present) continue;
        const int my_off = c.cand_off[srow + min(e_lane, N - 1)];
        const float bias_l = (bias && e_lane < N) ? bias[e_lane] : 0.f;
        for (; present; present &= present - 1) {
            const int i = __ffs(present) - 1;
            const uint32_t bits = __shfl_sync(FULL, my_bits, i);
            const int off = __shfl_sync(FULL, my_off, i);
            const float bl = __shfl_sync(FULL, bias_l, i);
            if ((bits >> lane) & 1u) {
                const int idx = off + __popc(bits & lt);
                const float l = bl + (idx < c.cap ? zc[idx] : 0.f);
                gt += (l > val) - (bias ? (bl > val) : 0);
                ge += (l >= val) - (bias ? (bl >= val) : 0);
            }
        }
    }
    sm_gt[warp][lane] = gt;
    sm_ge[warp][lane] = ge;
    __syncthreads();
    if (warp == 0) {
        int a = 0, b = 0;
        for (int k = 0; k < WARPS_PER_BLOCK; ++k) { a += sm_gt[k][lane]; b += sm_ge[k][lane]; }
        atomicAdd(counters + ((size_t)slot * 32 + lane) * 2, a);
        atomicAdd(counters + ((size_t)slot * 32 + lane) * 2 + 1, b);
    }
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_rank_cells_finalize(rl_graph g, rl_slots s, rl_cells c, rl_answers known, const float *__restrict__ bias,
                      const float *__restrict__ sorted_bias, const float *__restrict__ zc,
                      const int32_t *__restrict__ counters, int64_t *__restrict__ LH)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x;
    const int N = g.num_entities;
    const int q = s.slot_head[slot];
    const size_t srow = (size_t)slot * N;
    for (int b = warp; b < 32; b += WARPS_PER_BLOCK) {
        const int h = s.lane_h[slot * RL_LANES + b];
        const int t = s.lane_t[slot * RL_LANES + b];
        long long L = 0, H = 0;
        if (h >= 0 && t >= 0) {
            float val;
            int idx;
            if (!cell_logit(c, bias, zc, srow, t, b, val, idx)) { L = 1; H = (long long)N + 1; }   // mask[k,t] False (trainer.py:198-200)
            else {
                int gt = counters[((size_t)slot * 32 + b) * 2], ge = counters[((size_t)slot * 32 + b) * 2 + 1];
                if (bias) {                                     // entities whose logit is the bare bias: sorted table
                    int lo = 0, hi = N;                         // first index with sorted_bias > val
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_bias[mid] > val) hi = mid; else lo = mid + 1; }
                    gt += N - lo;
                    lo = 0; hi = N;                             // first index with sorted_bias >= val
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_bias[mid] >= val) hi = mid; else lo = mid + 1; }
                    ge += N - lo;
                }
                const int ki = find_key(known, (long long)q * N + h);
                const int a0 = ki >= 0 ? known.ptr[ki] : 0, a1 = ki >= 0 ? known.ptr[ki + 1] : 0;
                int fgt = 0, fge = 0;
                for (int a = a0 + lane; a < a1; a += 32) {      // flag False at the known answers (data.py:250-254)
                    float l;
                    int ix;
                    if (cell_logit(c, bias, zc, srow, known.ent[a], b, l, ix)) { fgt += l > val; fge += l >= val; }
                }
                fgt = warp_sumi(fgt);
                fge = warp_sumi(fge);
                L = (long long)(gt - fgt) + 1;
                H = (long long)(ge - fge) + 2;
            }
        }
        if (lane == 0) {
            LH[((size_t)slot * 32 + b) * 2] = L;
            LH[((size_t)slot * 32 + b) * 2 + 1] = H;
        }
    }
}

// ------------------------------------------------------------------------------------------
// cells <-> dense entity-major matrices (the RotatE entity feature is dense by nature; API forward())
// ------------------------------------------------------------------------------------------
// mode 0: Z[cell position] += zc     mode 1: Gc = G[cell position]
__global__ void __launch_bounds__(256)
k_cells_dense(int N, long long SN, rl_cells c, float *__restrict__ cellv, float *__restrict__ dense, int mode)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;      // (slot, entity)
    if (i >= SN) return;
    uint32_t bits = c.nzmask[i];
    int idx = c.cand_off[i];
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        if (idx < c.cap) {
            float *d = dense + (size_t)i * RL_LANES + b;
            if (mode == 0) *d += cellv[idx]; else cellv[idx] = *d;
        }
        ++idx;
    }
}

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
static int bad_cells(const rl_cells *c)
{
    return !c || !c->counters || !c->nzmask || !c->cand_off || !c->cell_key || !c->slot_ncell || c->cap <= 0;
}
static int bad_item_frontier(const rl_frontier *fr)
{
    return !fr || !fr->arena || !fr->items || !fr->items_sorted || !fr->item_cnt || !fr->item_off || !fr->bucket_cnt ||
           !fr->bucket_off || !fr->item_mask || !fr->item_mask_sorted || !fr->nzmask ||
           (fr->count_bits != 32 && fr->count_bits != 64);
}

extern "C" {

int rl_cells_build(const rl_graph *g, const rl_rules *r, const rl_slots *s, const rl_frontier *fr,
                   const rl_cells *c, void *stream)
{
    if (!g || !r || !s || bad_cells(c) || bad_item_frontier(fr)) return rl_fail(RL_ERR_ARG, "rl_cells_build: bad argument");
    if (c->nzmask != fr->nzmask) return rl_fail(RL_ERR_ARG, "rl_cells_build: cells and frontier must share nzmask");
    if (s->num_slots <= 0) return RL_OK;
    const int rc = rl_sort_items(g, s, fr, stream);               // groups items by entity, ORs their lane masks into nzmask
    if (rc != RL_OK) return rc;
    if (c->nnz_cap > 0 && (!c->nnz_off || !c->nz_val || !c->nz_cell || !c->nz_item))
        return rl_fail(RL_ERR_ARG, "rl_cells_build: nnz_cap > 0 needs the coordinate arrays");
    k_cell_scan<<<s->num_slots, 512, 0, (cudaStream_t)stream>>>(*g, *r, *s, *fr, *c);
    CHECK_LAUNCH("k_cell_scan");
    return RL_OK;
}

int rl_bias_stats(int32_t N, const float *bias, double *acc, void *stream)
{
    if (!bias || !acc || N <= 0) return rl_fail(RL_ERR_ARG, "rl_bias_stats: bad argument");
    k_bias_stats<<<1, 1024, 0, (cudaStream_t)stream>>>(N, bias, acc);
    CHECK_LAUNCH("k_bias_stats");
    return RL_OK;
}

int rl_predictor_cell_scores(const rl_graph *g, const rl_rules *r, const rl_slots *s, const rl_frontier *fr,
                             const rl_cells *c, const float *w, float *zc, void *stream)
{
    if (!g || !r || !s || !w || !zc || bad_cells(c) || bad_item_frontier(fr)) return rl_fail(RL_ERR_ARG, "rl_predictor_cell_scores: bad argument");
    if (s->num_slots <= 0) return RL_OK;
    const dim3 grid((g->rank_words + PC_WARPS - 1) / PC_WARPS, s->num_slots);
    cudaStream_t st = (cudaStream_t)stream;
    if (fr->count_bits == 32) k_pred_cells<uint32_t><<<grid, PC_WARPS * 32, 0, st>>>(*g, *r, *s, *fr, *c, w, zc);
    else k_pred_cells<unsigned long long><<<grid, PC_WARPS * 32, 0, st>>>(*g, *r, *s, *fr, *c, w, zc);
    CHECK_LAUNCH("k_pred_cells");
    return RL_OK;
}

int rl_cells_partial_floats(void) { return CELL_BLOCKS * 96; }

int rl_cells_softmax_ce(const rl_graph *g, const rl_slots *s, const rl_cells *c, const rl_answers *ans, float smoothing,
                        const float *bias, double *acc, const float *zc, int32_t n_groups, const int32_t *group_ptr,
                        float grad_scale, float *partial, float *stats, float *slot_sums, float *group_loss,
                        float *group_tsum, float *Gc, float *grad_bias, void *stream)
{
    if (!g || !s || !ans || !zc || !partial || !stats || !slot_sums || !group_loss || !group_tsum || bad_cells(c))
        return rl_fail(RL_ERR_ARG, "rl_cells_softmax_ce: null argument");
    if (bias && !acc) return rl_fail(RL_ERR_ARG, "rl_cells_softmax_ce: bias needs the rl_bias_stats accumulator");
    if (Gc && bias && !grad_bias) return rl_fail(RL_ERR_ARG, "rl_cells_softmax_ce: bias needs grad_bias");
    const int S = s->num_slots, N = g->num_entities;
    if (S <= 0) return RL_OK;
    if (n_groups <= 0 || n_groups > S || (!group_ptr && n_groups != S)) return rl_fail(RL_ERR_ARG, "rl_cells_softmax_ce: bad group table");
    cudaStream_t st = (cudaStream_t)stream;
    float *slot_lsum = slot_sums, *slot_tsum = slot_sums + S, *slot_invT = slot_sums + 2 * (size_t)S;
    k_ce_cells<<<S, CE_WARPS * 32, 0, st>>>(*g, *s, *c, *ans, smoothing, bias, acc, zc, stats, slot_lsum, slot_tsum);
    CHECK_LAUNCH("k_ce_cells");
    k_group_reduce<<<n_groups, 32, 0, st>>>(n_groups, group_ptr, slot_lsum, slot_tsum, group_loss, group_tsum, slot_invT, stats);
    CHECK_LAUNCH("k_group_reduce");
    if (!Gc) return RL_OK;
    if (bias) {                                                  // K of the rank-one bias gradient, accumulated by k_grad_cells
        cudaError_t e = cudaMemsetAsync(acc + 2, 0, sizeof(double), st);
        if (e != cudaSuccess) return rl_fail(RL_ERR_CUDA, "rl_cells_softmax_ce: memset", e);
    }
    k_grad_cells<<<dim3(CELL_BLOCKS, S), 256, 0, st>>>(*g, *c, bias, acc, zc, stats, grad_scale, Gc, grad_bias);
    CHECK_LAUNCH("k_grad_cells");
    k_grad_targets<<<S, CE_WARPS * 32, 0, st>>>(*g, *s, *c, *ans, smoothing, bias, zc, stats, slot_invT, grad_scale, Gc,
                                                 bias ? grad_bias : nullptr);
    CHECK_LAUNCH("k_grad_targets");
    if (bias) {
        k_bias_finish<<<(N + 255) / 256, 256, 0, st>>>(N, bias, acc, grad_bias);
        CHECK_LAUNCH("k_bias_finish");
    }
    return RL_OK;
}

int rl_predictor_cell_backward(const rl_graph *g, const rl_rules *r, const rl_slots *s, const rl_frontier *fr,
                               const rl_cells *c, const float *Gc, float *grad_w, void *stream)
{
    if (!g || !r || !s || !Gc || !grad_w || bad_cells(c) || bad_item_frontier(fr)) return rl_fail(RL_ERR_ARG, "rl_predictor_cell_backward: bad argument");
    if (s->num_slots <= 0) return RL_OK;
    const dim3 grid((g->rank_words + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, s->num_slots);
    cudaStream_t st = (cudaStream_t)stream;
    if (fr->count_bits == 32) k_pred_cells_bwd<uint32_t><<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(*g, *r, *s, *fr, *c, Gc, grad_w);
    else k_pred_cells_bwd<unsigned long long><<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(*g, *r, *s, *fr, *c, Gc, grad_w);
    CHECK_LAUNCH("k_pred_cells_bwd");
    return RL_OK;
}

int rl_predictor_nnz_backward(const rl_graph *g, const rl_rules *r, const rl_slots *s, const rl_frontier *fr,
                              const rl_cells *c, const float *Gc, float *grad_w, void *stream)
{
    if (!g || !r || !s || !Gc || !grad_w || bad_cells(c) || bad_item_frontier(fr)) return rl_fail(RL_ERR_ARG, "rl_predictor_nnz_backward: bad argument");
    if (c->nnz_cap <= 0 || !c->nz_val || !c->nz_cell || !c->nz_item) return rl_fail(RL_ERR_ARG, "rl_predictor_nnz_backward: no coordinate list");
    if (s->num_slots <= 0) return RL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    k_pred_nnz_bwd<<<148 * 8, 256, 0, st>>>(*r, *fr, *c, Gc, grad_w);
    CHECK_LAUNCH("k_pred_nnz_bwd");
    k_pred_zr_bwd<<<s->num_slots, 32, 0, st>>>(*g, *r, *s, *c, Gc, grad_w);
    CHECK_LAUNCH("k_pred_zr_bwd");
    return RL_OK;
}

int rl_cells_rank(const rl_graph *g, const rl_slots *s, const rl_cells *c, const rl_answers *known, const float *bias,
                  const float *sorted_bias, const float *zc, int32_t *counters, int64_t *LH, void *stream)
{
    if (!g || !s || !known || !zc || !counters || !LH || bad_cells(c)) return rl_fail(RL_ERR_ARG, "rl_cells_rank: null argument");
    if (bias && !sorted_bias) return rl_fail(RL_ERR_ARG, "rl_cells_rank: bias needs its ascending-sorted copy");
    const int S = s->num_slots;
    if (S <= 0) return RL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counters, 0, (size_t)S * 64 * sizeof(int32_t), st);
    if (e != cudaSuccess) return rl_fail(RL_ERR_CUDA, "rl_cells_rank: memset", e);
    k_rank_cells<<<dim3(CELL_BLOCKS, S), WARPS_PER_BLOCK * 32, 0, st>>>(*g, *s, *c, bias, zc, counters);
    CHECK_LAUNCH("k_rank_cells");
    k_rank_cells_finalize<<<S, WARPS_PER_BLOCK * 32, 0, st>>>(*g, *s, *c, *known, bias, sorted_bias, zc, counters, LH);
    CHECK_LAUNCH("k_rank_cells_finalize");
    return RL_OK;
}

int rl_cells_add_to_dense(const rl_graph *g, const rl_slots *s, const rl_cells *c, const float *zc, float *Z, void *stream)
{
    if (!g || !s || !zc || !Z || bad_cells(c)) return rl_fail(RL_ERR_ARG, "rl_cells_add_to_dense: null argument");
    const long long SN = (long long)s->num_slots * g->num_entities;
    if (SN <= 0) return RL_OK;
    k_cells_dense<<<(unsigned)((SN + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g->num_entities, SN, *c, const_cast<float *>(zc), Z, 0);
    CHECK_LAUNCH("k_cells_dense");
    return RL_OK;
}

int rl_cells_gather_dense(const rl_graph *g, const rl_slots *s, const rl_cells *c, const float *G, float *Gc, void *stream)
{
    if (!g || !s || !G || !Gc || bad_cells(c)) return rl_fail(RL_ERR_ARG, "rl_cells_gather_dense: null argument");
    const long long SN = (long long)s->num_slots * g->num_entities;
    if (SN <= 0) return RL_OK;
    k_cells_dense<<<(unsigned)((SN + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g->num_entities, SN, *c, Gc, const_cast<float *>(G), 1);
    CHECK_LAUNCH("k_cells_dense");
    return RL_OK;
}

}  // extern "C"
