// rnnlogic_b200 -- the PNA aggregator of PredictorPlus (FuncToNode, src/layers.py:89-126; the shipped WN18RR config)
// on candidate cells: per-cell statistics from the item list, the scaler / Linear(12H,H) front, and their backward
// into the rule embeddings and the 16 x 192 weight.  H = 16.  The LayerNorm -> ReLU -> MLP tail is rl_tail_tc.cu with
// front_done = 1.
//
//   degree = sum_rule count + 1                                   (layers.py:92)
//   mean = S1 / degree, S1 = sum count * emb;  sq_mean = S2 / degree, S2 = sum count * emb^2
//   mn / mx = min / max of emb over the rules with count != 0     (layers.py:96-99)
//   std = sqrt(clamp(sq_mean - mean^2, 1e-6))
//   s = log(degree) / mean over the query's cells of log(degree);  scalers = [1, s, 1/clamp(s)]   (layers.py:109-116)
//   y = W [mean, mn, mx, std] (x) scalers + b                     (layers.py:118-124: Linear(12H, H))
//
// The statistics are accumulated by one thread per (item, quarter of the hidden vector) with vector atomics; min / max
// carry the rule in the low word of a 64-bit key (order-preserving float key << 32 | rule tag), so the arg rule of the
// backward -- the FIRST rule in rule-file order on ties, like torch.min / max -- falls out of the same atomic.
#include "rl_device.cuh"
#include "rl_umma.cuh"

#define PH 16                 // hidden_dim
#define PF 64                 // features per cell: mean | min | max | std
#define PU 192                // inputs of the Linear: feature f, scaler t -> f*3 + t
#define PNA_BLOCKS 16

__device__ __forceinline__ unsigned pkey(float v)
{
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float pkey_inv(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ void pna_add(const rl_pna &p, long long cell, int part, float v, const float4 e4, int rule)
{
    atomicAdd(reinterpret_cast<float4 *>(p.s1 + cell * PH) + part, make_float4(v * e4.x, v * e4.y, v * e4.z, v * e4.w));
    atomicAdd(reinterpret_cast<float4 *>(p.s2 + cell * PH) + part,
              make_float4(v * (e4.x * e4.x), v * (e4.y * e4.y), v * (e4.z * e4.z), v * (e4.w * e4.w)));
    if (part == 0) atomicAdd(p.deg + cell, v);
    const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
    const unsigned long long lo_min = (unsigned)rule, lo_max = 0xffffffffu - (unsigned)rule;   // ties: the first rule wins both
    // look before the atomic: a key only ever moves towards its extreme, so a (possibly stale) L2 read that already beats
    // this rule means the atomic cannot change anything -- most contributions of a well-connected cell skip it
    unsigned long long cur_mn[4], cur_mx[4];
#pragma unroll
    for (int u2 = 0; u2 < 2; ++u2) {
        const ulonglong2 a = __ldcg(reinterpret_cast<const ulonglong2 *>(p.mnk + cell * PH + part * 4) + u2);
        const ulonglong2 b = __ldcg(reinterpret_cast<const ulonglong2 *>(p.mxk + cell * PH + part * 4) + u2);
        cur_mn[2 * u2] = a.x; cur_mn[2 * u2 + 1] = a.y; cur_mx[2 * u2] = b.x; cur_mx[2 * u2 + 1] = b.y;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const unsigned long long k = (unsigned long long)pkey(ev[u]) << 32;
        if ((k | lo_min) < cur_mn[u]) atomicMin(p.mnk + cell * PH + part * 4 + u, k | lo_min);
        if ((k | lo_max) > cur_mx[u]) atomicMax(p.mxk + cell * PH + part * 4 + u, k | lo_max);
    }
}

template <typename CT>
__global__ void __launch_bounds__(256)
k_pna_item_stats(rl_graph g, rl_rules r, rl_slots s, rl_frontier fr, rl_cells c, const float *__restrict__ emb, rl_pna p)
{
    const int slot = blockIdx.y;
    const int n = fr.item_cnt[slot];
    const long long ib = fr.item_off[slot];
    const int4 *items = reinterpret_cast<const int4 *>(fr.items) + ib;
    const uint32_t *masks = fr.item_mask + ib;
    const CT *arena = reinterpret_cast<const CT *>(fr.arena) + (size_t)s.arena_off[slot] * RL_LANES;
    const size_t srow = (size_t)slot * g.num_entities;
    const int part = threadIdx.x & 3;
    for (int i = (blockIdx.x * 256 + threadIdx.x) >> 2; i < n; i += PNA_BLOCKS * 64) {
        const uint32_t m0 = __ldg(masks + i);
        if (!m0) continue;
        const int4 it = __ldg(items + i);                         // {row, first rule end, entity, rule ends}
        const uint32_t bits = c.nzmask[srow + it.z];
        const int off = c.cand_off[srow + it.z];
        const CT *row = arena + (size_t)it.x * RL_LANES;
        for (int t = it.y; t < it.y + it.w; ++t) {
            const int rule = __ldg(r.node_term_rule + t);
            const float4 e4 = __ldg(reinterpret_cast<const float4 *>(emb + (size_t)rule * PH) + part);
            for (uint32_t m = m0; m; m &= m - 1) {
                const int b = __ffs(m) - 1;
                const long long cell = off + __popc(bits & ((1u << b) - 1u));
                if (cell < c.cap) pna_add(p, cell, part, (float)row[b], e4, rule);
            }
        }
    }
}

// empty-body rules: count = one_hot(h) at the cell (h_b, b)
__global__ void __launch_bounds__(128)
k_pna_zr_stats(rl_graph g, rl_rules r, rl_slots s, rl_cells c, const float *__restrict__ emb, rl_pna p)
{
    const int slot = blockIdx.x, b = threadIdx.x >> 2, part = threadIdx.x & 3;
    const int q = s.slot_head[slot];
    const int z0 = r.zr_ptr[q], z1 = r.zr_ptr[q + 1];
    if (z1 <= z0) return;
    const int h = s.lane_h[slot * RL_LANES + b];
    if (h < 0) return;
    const size_t srow = (size_t)slot * g.num_entities;
    const uint32_t bits = c.nzmask[srow + h];
    const long long cell = c.cand_off[srow + h] + __popc(bits & ((1u << b) - 1u));
    if (cell >= c.cap) return;
    for (int t = z0; t < z1; ++t) {
        const int rule = r.zr_rule[t];
        pna_add(p, cell, part, 1.f, __ldg(reinterpret_cast<const float4 *>(emb + (size_t)rule * PH) + part), rule);
    }
}

// per-query sum and count of log(degree) over the query's cells (layers.py:109-113); a block's 256 consecutive cells belong
// to one or two slots (the first cell's and the last cell's): combined in shared memory first, as in k_cell_sum
__global__ void __launch_bounds__(256)
k_pna_qscale(rl_cells c, rl_pna p, float *__restrict__ qlog, float *__restrict__ qn)
{
    __shared__ float sl[64], sn[64];
    const int n = min(c.counters[0], c.cap);
    for (int base = blockIdx.x * 256; base < n; base += gridDim.x * 256) {
        if (threadIdx.x < 64) { sl[threadIdx.x] = 0.f; sn[threadIdx.x] = 0.f; }
        __syncthreads();
        const int ka = c.cell_key[base] & ~31, kb = c.cell_key[min(base + 255, n - 1)] & ~31;
        const int i = base + threadIdx.x;
        if (i < n) {
            const int key = c.cell_key[i], ks = key & ~31;
            const float lg = logf(p.deg[i] + 1.f);
            if (ks == ka) { atomicAdd(sl + (key & 31), lg); atomicAdd(sn + (key & 31), 1.f); }
            else if (ks == kb) { atomicAdd(sl + 32 + (key & 31), lg); atomicAdd(sn + 32 + (key & 31), 1.f); }
            else { atomicAdd(qlog + key, lg); atomicAdd(qn + key, 1.f); }
        }
        __syncthreads();
        if (threadIdx.x < 64 && sn[threadIdx.x] != 0.f) {
            const int key = (threadIdx.x < 32 ? ka : kb) + (threadIdx.x & 31);
            atomicAdd(qlog + key, sl[threadIdx.x]);
            atomicAdd(qn + key, sn[threadIdx.x]);
        }
        __syncthreads();
    }
}

// features of one cell into registers; returns the scaler s
__device__ __forceinline__ float pna_features(const rl_cells &c, const rl_pna &p, long long cell, const float *__restrict__ qlog,
                                              const float *__restrict__ qn, float (&feat)[PF])
{
    const float deg1 = p.deg[cell] + 1.f;
    const float dcl = fmaxf(deg1, 1e-6f);
    // 16-byte loads: a cell's statistics are 64-byte (sums) and 128-byte (min / max keys) rows
#pragma unroll
    for (int h4 = 0; h4 < PH / 4; ++h4) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p.s1 + cell * PH) + h4);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(p.s2 + cell * PH) + h4);
        const float s1v[4] = {a.x, a.y, a.z, a.w}, s2v[4] = {b.x, b.y, b.z, b.w};
        unsigned long long mn[4], mx[4];
#pragma unroll
        for (int u2 = 0; u2 < 2; ++u2) {
            const ulonglong2 q = __ldg(reinterpret_cast<const ulonglong2 *>(p.mnk + cell * PH) + 2 * h4 + u2);
            const ulonglong2 r = __ldg(reinterpret_cast<const ulonglong2 *>(p.mxk + cell * PH) + 2 * h4 + u2);
            mn[2 * u2] = q.x; mn[2 * u2 + 1] = q.y; mx[2 * u2] = r.x; mx[2 * u2 + 1] = r.y;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int h = 4 * h4 + u;
            const float mean = s1v[u] / dcl;
            const float sqm = s2v[u] / dcl;
            feat[h] = mean;
            feat[PH + h] = pkey_inv((unsigned)(mn[u] >> 32));
            feat[2 * PH + h] = pkey_inv((unsigned)(mx[u] >> 32));
            feat[3 * PH + h] = sqrtf(fmaxf(sqm - mean * mean, 1e-6f));
        }
    }
    const int key = c.cell_key[cell];
    const float msc = qlog[key] / fmaxf(qn[key], 1e-6f);
    return logf(deg1) / fmaxf(msc, 1e-6f);
}

// ---- the Linear(12H,H) front on the tensor cores (tcgen05, TMEM accumulators; helpers in rl_umma.cuh) ----
// y = b + W [f (x) (1, s, 1/s)]: with W3[t*16+i][f] = W[i][f*3+t] the three scaler blocks are ONE product
// P[cell][48] = FEAT[cell][64] W3^T and y_i = b_i + P_i + s P_{16+i} + P_{32+i} / s.  A tile is 128 cells = the M of the MMA;
// both operands are general fp32, so the product is 3xTF32 (hi/lo pieces, fp32 accumulate): 24 MMAs of K = 8, N = 48.
#define PT 128
#define PCH (PT * 16)                         // bytes of one 16-byte chunk column of a 128-row tile
#define PF_AHI 0                              // FEAT hi: 16 chunks of 4 features
#define PF_ALO (PF_AHI + 16 * PCH)
#define PF_BHI (PF_ALO + 16 * PCH)            // W3 hi: [48 rows][16 chunks]
#define PF_BLO (PF_BHI + 16 * 48 * 16)
#define PF_BAR (PF_BLO + 16 * 48 * 16)
#define PF_END (PF_BAR + 16)

__global__ void __launch_bounds__(PT, 2)
k_pna_front_fwd(rl_cells c, rl_pna p, const float *__restrict__ qlog, const float *__restrict__ qn,
                const float *__restrict__ W, const float *__restrict__ bvec, float *__restrict__ Y,
                float *__restrict__ FEAT, float *__restrict__ SC)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const long long C = min(c.counters[0], c.cap);
    const long long tiles = (C + PT - 1) / PT;
    if ((long long)blockIdx.x >= tiles) return;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + PF_BAR);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(smem + PF_BAR + 8);
    for (int x = tid; x < 48 * 16; x += PT) {                    // (row n = t*16+i, chunk of 4 features)
        const int n = x % 48, ch = x / 48, t = n >> 4, i = n & 15;
        float4 hi, lo;
        umma::split_tf32(__ldg(W + i * PU + (4 * ch) * 3 + t), hi.x, lo.x);
        umma::split_tf32(__ldg(W + i * PU + (4 * ch + 1) * 3 + t), hi.y, lo.y);
        umma::split_tf32(__ldg(W + i * PU + (4 * ch + 2) * 3 + t), hi.z, lo.z);
        umma::split_tf32(__ldg(W + i * PU + (4 * ch + 3) * 3 + t), hi.w, lo.w);
        *reinterpret_cast<float4 *>(smem + PF_BHI + ch * (48 * 16) + n * 16) = hi;
        *reinterpret_cast<float4 *>(smem + PF_BLO + ch * (48 * 16) + n * 16) = lo;
    }
    if (tid == 0) umma::mbar_init(bar, 1);
    if (warp == 0) umma::tmem_alloc(tslot, 64);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tacc = *tslot, trow = tacc + ((uint32_t)(warp * 32) << 16);
    const uint32_t s_ahi = umma::smem_u32(smem + PF_AHI), s_alo = umma::smem_u32(smem + PF_ALO);
    const uint32_t s_bhi = umma::smem_u32(smem + PF_BHI), s_blo = umma::smem_u32(smem + PF_BLO);
    constexpr uint32_t ID = umma::idesc(UMMA_FMT_TF32, PT, 48, false, false);
    float bv[PH];
#pragma unroll
    for (int i = 0; i < PH; ++i) bv[i] = __ldg(bvec + i);
    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long cell = tile * PT + tid;
        const bool live = cell < C;
        float sc = 1.f;
        {
            float feat[PF];
            if (live) sc = pna_features(c, p, cell, qlog, qn, feat);
            else {
#pragma unroll
                for (int f = 0; f < PF; ++f) feat[f] = 0.f;
            }
#pragma unroll
            for (int ch = 0; ch < PF / 4; ++ch) {
                float4 hi, lo;
                umma::split_tf32(feat[4 * ch], hi.x, lo.x); umma::split_tf32(feat[4 * ch + 1], hi.y, lo.y);
                umma::split_tf32(feat[4 * ch + 2], hi.z, lo.z); umma::split_tf32(feat[4 * ch + 3], hi.w, lo.w);
                *reinterpret_cast<float4 *>(smem + PF_AHI + ch * PCH + tid * 16) = hi;
                *reinterpret_cast<float4 *>(smem + PF_ALO + ch * PCH + tid * 16) = lo;
                if (live) reinterpret_cast<float4 *>(FEAT + cell * PF)[ch] = make_float4(feat[4 * ch], feat[4 * ch + 1], feat[4 * ch + 2], feat[4 * ch + 3]);
            }
            if (live) SC[cell] = sc;
        }
        umma::fence_smem_to_async();
        umma::fence_before();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after();
#pragma unroll
            for (int pr = 0; pr < 3; ++pr) {                     // lo*hi, hi*lo, hi*hi
                const uint32_t sa = pr == 0 ? s_alo : s_ahi, sb = pr == 1 ? s_blo : s_bhi;
#pragma unroll
                for (int k = 0; k < PF / 8; ++k)
                    umma::mma_tf32(tacc, umma::desc(sa + k * 2 * PCH, PCH, 128), umma::desc(sb + k * 2 * (48 * 16), 48 * 16, 128), ID, (pr | k) != 0);
            }
            umma::commit(bar);
        }
        umma::mbar_wait(bar, phase);
        phase ^= 1u;
        umma::fence_after();
        float d0[16], d1[16], d2[16];
        umma::tmem_ld16(trow, d0);
        umma::tmem_ld16(trow + 16, d1);
        umma::tmem_ld16(trow + 32, d2);
        if (live) {
            const float isc = 1.f / fmaxf(sc, 1e-6f);
            float4 *yo = reinterpret_cast<float4 *>(Y + cell * PH);
#pragma unroll
            for (int i4 = 0; i4 < PH / 4; ++i4) {
                float y[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = 4 * i4 + u;
                    y[u] = fmaf(isc, d2[i], fmaf(sc, d1[i], d0[i] + bv[i]));
                }
                yo[i4] = make_float4(y[0], y[1], y[2], y[3]);
            }
        }
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tacc, 64);
}

// backward of the front: dfeat[cell][64] = [dy, s dy, dy / s][cell][48] W3, again 3xTF32 (18 MMAs of K = 8, N = 64);
// then dstat[cell] = [dS1 16 | dS2 16 | dMN 16 | dMX 16] per thread.
#define PB_AHI 0                              // [dy, s dy, dy/s] hi: 12 chunks of 4
#define PB_ALO (PB_AHI + 12 * PCH)
#define PB_BHI (PB_ALO + 12 * PCH)            // W3^T hi: [64 rows f][12 chunks of n]
#define PB_BLO (PB_BHI + 12 * 64 * 16)
#define PB_BAR (PB_BLO + 12 * 64 * 16)
#define PB_END (PB_BAR + 16)

__global__ void __launch_bounds__(PT, 2)
k_pna_front_bwd(rl_cells c, rl_pna p, const float *__restrict__ W, const float *__restrict__ dY,
                const float *__restrict__ FEAT, const float *__restrict__ SC, float *__restrict__ dstat)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const long long C = min(c.counters[0], c.cap);
    const long long tiles = (C + PT - 1) / PT;
    if ((long long)blockIdx.x >= tiles) return;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + PB_BAR);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(smem + PB_BAR + 8);
    for (int x = tid; x < 64 * 12; x += PT) {                    // (row f, chunk of 4 n = t*16+i)
        const int f = x & 63, ch = x >> 6, t = ch >> 2, i0 = (ch & 3) * 4;
        float4 hi, lo;
        umma::split_tf32(__ldg(W + (i0) * PU + f * 3 + t), hi.x, lo.x);
        umma::split_tf32(__ldg(W + (i0 + 1) * PU + f * 3 + t), hi.y, lo.y);
        umma::split_tf32(__ldg(W + (i0 + 2) * PU + f * 3 + t), hi.z, lo.z);
        umma::split_tf32(__ldg(W + (i0 + 3) * PU + f * 3 + t), hi.w, lo.w);
        *reinterpret_cast<float4 *>(smem + PB_BHI + ch * (64 * 16) + f * 16) = hi;
        *reinterpret_cast<float4 *>(smem + PB_BLO + ch * (64 * 16) + f * 16) = lo;
    }
    if (tid == 0) umma::mbar_init(bar, 1);
    if (warp == 0) umma::tmem_alloc(tslot, 64);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tacc = *tslot, trow = tacc + ((uint32_t)(warp * 32) << 16);
    const uint32_t s_ahi = umma::smem_u32(smem + PB_AHI), s_alo = umma::smem_u32(smem + PB_ALO);
    const uint32_t s_bhi = umma::smem_u32(smem + PB_BHI), s_blo = umma::smem_u32(smem + PB_BLO);
    constexpr uint32_t ID = umma::idesc(UMMA_FMT_TF32, PT, 64, false, false);
    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long cell = tile * PT + tid;
        const bool live = cell < C;
        {
            const float sc = live ? SC[cell] : 1.f, isc = 1.f / fmaxf(sc, 1e-6f);
#pragma unroll
            for (int i4 = 0; i4 < PH / 4; ++i4) {
                const float4 v = live ? __ldg(reinterpret_cast<const float4 *>(dY + cell * PH) + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float dv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const float m = t == 0 ? 1.f : (t == 1 ? sc : isc);
                    float4 hi, lo;
                    umma::split_tf32(dv[0] * m, hi.x, lo.x); umma::split_tf32(dv[1] * m, hi.y, lo.y);
                    umma::split_tf32(dv[2] * m, hi.z, lo.z); umma::split_tf32(dv[3] * m, hi.w, lo.w);
                    *reinterpret_cast<float4 *>(smem + PB_AHI + (t * 4 + i4) * PCH + tid * 16) = hi;
                    *reinterpret_cast<float4 *>(smem + PB_ALO + (t * 4 + i4) * PCH + tid * 16) = lo;
                }
            }
        }
        umma::fence_smem_to_async();
        umma::fence_before();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after();
#pragma unroll
            for (int pr = 0; pr < 3; ++pr) {
                const uint32_t sa = pr == 0 ? s_alo : s_ahi, sb = pr == 1 ? s_blo : s_bhi;
#pragma unroll
                for (int k = 0; k < 48 / 8; ++k)
                    umma::mma_tf32(tacc, umma::desc(sa + k * 2 * PCH, PCH, 128), umma::desc(sb + k * 2 * (64 * 16), 64 * 16, 128), ID, (pr | k) != 0);
            }
            umma::commit(bar);
        }
        umma::mbar_wait(bar, phase);
        phase ^= 1u;
        umma::fence_after();
        float dmean[16], dmn[16], dmx[16], dsd[16];              // gradient of feature (kind, unit h): columns kind*16 + h
        umma::tmem_ld16(trow, dmean);
        umma::tmem_ld16(trow + 16, dmn);
        umma::tmem_ld16(trow + 32, dmx);
        umma::tmem_ld16(trow + 48, dsd);
        if (live) {
            const float deg1 = fmaxf(p.deg[cell] + 1.f, 1e-6f);
            float4 *out = reinterpret_cast<float4 *>(dstat + cell * PF);
#pragma unroll
            for (int h4 = 0; h4 < PH / 4; ++h4) {
                const float4 mean4 = __ldg(reinterpret_cast<const float4 *>(FEAT + cell * PF) + h4);
                const float4 std4 = __ldg(reinterpret_cast<const float4 *>(FEAT + cell * PF + 3 * PH) + h4);
                const float4 s24 = __ldg(reinterpret_cast<const float4 *>(p.s2 + cell * PH) + h4);
                const float mean[4] = {mean4.x, mean4.y, mean4.z, mean4.w}, stdv[4] = {std4.x, std4.y, std4.z, std4.w},
                            s2v[4] = {s24.x, s24.y, s24.z, s24.w};
                float o1[4], o2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int h = 4 * h4 + u;
                    const float var = s2v[u] / deg1 - mean[u] * mean[u];             // the forward's expression, bit for bit
                    const float dv = var >= 1e-6f ? dsd[h] / (2.f * stdv[u]) : 0.f;  // clamp(min=1e-6) passes the gradient where v >= 1e-6
                    o1[u] = (dmean[h] - 2.f * mean[u] * dv) / deg1;                  // dS1
                    o2[u] = dv / deg1;                                               // dS2
                }
                out[h4] = make_float4(o1[0], o1[1], o1[2], o1[3]);
                out[PH / 4 + h4] = make_float4(o2[0], o2[1], o2[2], o2[3]);
                out[2 * (PH / 4) + h4] = make_float4(dmn[4 * h4], dmn[4 * h4 + 1], dmn[4 * h4 + 2], dmn[4 * h4 + 3]);
                out[3 * (PH / 4) + h4] = make_float4(dmx[4 * h4], dmx[4 * h4 + 1], dmx[4 * h4 + 2], dmx[4 * h4 + 3]);
            }
        }
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tacc, 64);
}

// gW[i][f*3+t] += sum_cells dy_i * feat_f * scaler_t  -- a product over the CELLS: gW3[n][f] = sum_c DY3[c][n] FEAT[c][f]
// (n = t*16+i).  On the tensor cores with K = cells: both operands are K-major tiles whose rows are the n / f indices,
// so a thread (= cell) writes its values as one 4-byte COLUMN of the tile; a row count of 129 (one pad row) spreads the
// eight 16-byte chunks of a warp over all 32 banks.  hi / lo pieces are stacked along M and N:
//     A rows: [DY3 hi 48 | 0 x16 | DY3 lo 48 | 0 x16]      B rows: [FEAT hi 64 | FEAT lo 64]
// and D[128][128] accumulates in TMEM over the whole kernel; gW3 = D[n][f] + D[n][64+f] + D[64+n][f] (3xTF32).
#define PW_ROWS 129
#define PW_CH (PW_ROWS * 16)                  // bytes of one chunk (4 cells) of a tile
#define PW_A 0
#define PW_B (PW_A + 32 * PW_CH)
#define PW_BAR (PW_B + 32 * PW_CH)
#define PW_END (PW_BAR + 16)

__global__ void __launch_bounds__(PT, 1)
k_pna_w_grad(rl_cells c, const float *__restrict__ dY, const float *__restrict__ FEAT, const float *__restrict__ SC,
             float *__restrict__ gW)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const long long C = min(c.counters[0], c.cap);
    const long long tiles = (C + PT - 1) / PT;
    const long long t0 = tiles * blockIdx.x / gridDim.x, t1 = tiles * (blockIdx.x + 1) / gridDim.x;
    if (t0 >= t1) return;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + PW_BAR);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(smem + PW_BAR + 8);
    for (int x = tid; x < 32 * 32; x += PT) {                    // the zero rows of A: 48..63 and 112..127, every chunk
        const int ch = x >> 5, r = x & 31;
        const int row = r < 16 ? 48 + r : 96 + r;
        *reinterpret_cast<float4 *>(smem + PW_A + ch * PW_CH + row * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) umma::mbar_init(bar, 1);
    if (warp == 0) umma::tmem_alloc(tslot, 128);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tacc = *tslot, trow = tacc + ((uint32_t)(warp * 32) << 16);
    const uint32_t s_a = umma::smem_u32(smem + PW_A), s_b = umma::smem_u32(smem + PW_B);
    constexpr uint32_t ID = umma::idesc(UMMA_FMT_TF32, 128, 128, false, false);
    float *colA = reinterpret_cast<float *>(smem + PW_A + (tid >> 2) * PW_CH + (tid & 3) * 4);     // this cell's column
    float *colB = reinterpret_cast<float *>(smem + PW_B + (tid >> 2) * PW_CH + (tid & 3) * 4);
    uint32_t phase = 0;
    for (long long tile = t0; tile < t1; ++tile) {
        const long long cell = tile * PT + tid;
        const bool live = cell < C;
        const float sc = live ? SC[cell] : 0.f, isc = live ? 1.f / fmaxf(sc, 1e-6f) : 0.f;
#pragma unroll
        for (int i4 = 0; i4 < PH / 4; ++i4) {
            const float4 v = live ? __ldg(reinterpret_cast<const float4 *>(dY + cell * PH) + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float dv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = 4 * i4 + u;
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    float hi, lo;
                    umma::split_tf32(dv[u] * (t == 0 ? 1.f : (t == 1 ? sc : isc)), hi, lo);
                    colA[(t * 16 + i) * 4] = hi;                 // row stride 16 bytes = 4 floats
                    colA[(64 + t * 16 + i) * 4] = lo;
                }
            }
        }
#pragma unroll
        for (int f4 = 0; f4 < PF / 4; ++f4) {
            const float4 v = live ? __ldg(reinterpret_cast<const float4 *>(FEAT + cell * PF) + f4) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float fv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float hi, lo;
                umma::split_tf32(fv[u], hi, lo);
                colB[(4 * f4 + u) * 4] = hi;
                colB[(64 + 4 * f4 + u) * 4] = lo;
            }
        }
        umma::fence_smem_to_async();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after();
#pragma unroll
            for (int k = 0; k < PT / 8; ++k)                     // K = 8 cells = two chunks per MMA
                umma::mma_tf32(tacc, umma::desc(s_a + k * 2 * PW_CH, PW_CH, 128), umma::desc(s_b + k * 2 * PW_CH, PW_CH, 128), ID,
                               tile != t0 || k != 0);
            umma::commit(bar);
        }
        umma::mbar_wait(bar, phase);                             // the operand tiles are free again
        phase ^= 1u;
    }
    umma::fence_after();
    {   // lane m of the accumulator: m < 48 -> hi rows (hi*hi + hi*lo), 64 <= m < 112 -> lo rows (lo*hi)
        const int m = tid;
        const bool hi_row = m < 48, lo_row = m >= 64 && m < 112;
        const int n = hi_row ? m : m - 64, t = n >> 4, i = n & 15;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float a[32], b[32];
            umma::tmem_ld32(trow + q * 32, a);                   // columns f = 32q .. (FEAT hi)
            umma::tmem_ld32(trow + 64 + q * 32, b);              // columns 64 + f  (FEAT lo)
            if (hi_row || lo_row) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float v = hi_row ? a[j] + b[j] : a[j];
                    if (v != 0.f) atomicAdd(gW + i * PU + (32 * q + j) * 3 + t, v);
                }
            }
        }
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tacc, 128);
}

// backward into the rule embeddings: gEmb[rule] += count * dS1 + 2 count emb * dS2 (+ dMN / dMX at the arg rules)
template <typename CT>
__global__ void __launch_bounds__(256)
k_pna_item_bwd(rl_graph g, rl_rules r, rl_slots s, rl_frontier fr, rl_cells c, const float *__restrict__ emb, rl_pna p,
               const float *__restrict__ dstat, float *__restrict__ gEmb)
{
    const int slot = blockIdx.y;
    const int n = fr.item_cnt[slot];
    const long long ib = fr.item_off[slot];
    const int4 *items = reinterpret_cast<const int4 *>(fr.items) + ib;
    const uint32_t *masks = fr.item_mask + ib;
    const CT *arena = reinterpret_cast<const CT *>(fr.arena) + (size_t)s.arena_off[slot] * RL_LANES;
    const size_t srow = (size_t)slot * g.num_entities;
    const int part = threadIdx.x & 3;
    for (int i = (blockIdx.x * 256 + threadIdx.x) >> 2; i < n; i += PNA_BLOCKS * 64) {
        const uint32_t m0 = __ldg(masks + i);
        if (!m0) continue;
        const int4 it = __ldg(items + i);
        const uint32_t bits = c.nzmask[srow + it.z];
        const int off = c.cand_off[srow + it.z];
        const CT *row = arena + (size_t)it.x * RL_LANES;
        for (int t = it.y; t < it.y + it.w; ++t) {
            const int rule = __ldg(r.node_term_rule + t);
            const float4 e4 = __ldg(reinterpret_cast<const float4 *>(emb + (size_t)rule * PH) + part);
            const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
            float a[4] = {0.f, 0.f, 0.f, 0.f};
            for (uint32_t m = m0; m; m &= m - 1) {
                const int b = __ffs(m) - 1;
                const long long cell = off + __popc(bits & ((1u << b) - 1u));
                if (cell >= c.cap) continue;
                const float v = (float)row[b];
                const float *ds = dstat + cell * PF + part * 4;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    a[u] += v * ds[u] + 2.f * v * ev[u] * ds[PH + u];
                    if ((unsigned)p.mnk[cell * PH + part * 4 + u] == (unsigned)rule) a[u] += ds[2 * PH + u];
                    if ((unsigned)p.mxk[cell * PH + part * 4 + u] == 0xffffffffu - (unsigned)rule) a[u] += ds[3 * PH + u];
                }
            }
            atomicAdd(reinterpret_cast<float4 *>(gEmb + (size_t)rule * PH) + part, make_float4(a[0], a[1], a[2], a[3]));
        }
    }
}

__global__ void __launch_bounds__(128)
k_pna_zr_bwd(rl_graph g, rl_rules r, rl_slots s, rl_cells c, const float *__restrict__ emb, rl_pna p,
             const float *__restrict__ dstat, float *__restrict__ gEmb)
{
    const int slot = blockIdx.x, b = threadIdx.x >> 2, part = threadIdx.x & 3;
    const int q = s.slot_head[slot];
    const int z0 = r.zr_ptr[q], z1 = r.zr_ptr[q + 1];
    if (z1 <= z0) return;
    const int h = s.lane_h[slot * RL_LANES + b];
    if (h < 0) return;
    const size_t srow = (size_t)slot * g.num_entities;
    const uint32_t bits = c.nzmask[srow + h];
    const long long cell = c.cand_off[srow + h] + __popc(bits & ((1u << b) - 1u));
    if (cell >= c.cap) return;
    const float *ds = dstat + cell * PF + part * 4;
    for (int t = z0; t < z1; ++t) {
        const int rule = r.zr_rule[t];
        const float4 e4 = __ldg(reinterpret_cast<const float4 *>(emb + (size_t)rule * PH) + part);
        const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
        float a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a[u] = ds[u] + 2.f * ev[u] * ds[PH + u];
            if ((unsigned)p.mnk[cell * PH + part * 4 + u] == (unsigned)rule) a[u] += ds[2 * PH + u];
            if ((unsigned)p.mxk[cell * PH + part * 4 + u] == 0xffffffffu - (unsigned)rule) a[u] += ds[3 * PH + u];
        }
        atomicAdd(reinterpret_cast<float4 *>(gEmb + (size_t)rule * PH) + part, make_float4(a[0], a[1], a[2], a[3]));
    }
}

static int bad_pna(const rl_cells *c, const rl_pna *p)
{
    return !c || !c->counters || !c->nzmask || !c->cand_off || !c->cell_key || c->cap <= 0 || !p || !p->s1 || !p->s2 || !p->deg ||
           !p->mnk || !p->mxk;
}
static int bad_items(const rl_frontier *fr)
{
    return !fr || !fr->arena || !fr->items || !fr->item_off || !fr->item_cnt || !fr->item_mask ||
           (fr->count_bits != 32 && fr->count_bits != 64);
}
static int cell_grid(const rl_cells *c)
{
    const long long b = ((long long)c->cap + 255) / 256;
    return (int)(b < 148 * 8 ? b : 148 * 8);
}

static int tile_grid(const rl_cells *c)
{
    const long long b = ((long long)c->cap + PT - 1) / PT;
    return (int)(b < 148 * 2 ? (b > 0 ? b : 1) : 148 * 2);
}

extern "C" {

int rl_pna_item_stats(const rl_graph *g, const rl_rules *r, const rl_slots *s, const rl_frontier *fr, const rl_cells *c,
                      const float *emb, const rl_pna *p, void *stream)
{
    if (!g || !r || !s || !emb || bad_pna(c, p) || bad_items(fr)) return rl_fail(RL_ERR_ARG, "rl_pna_item_stats: bad argument");
    if (s->num_slots <= 0) return RL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)c->cap;
    if (cudaMemsetAsync(p->s1, 0, n * PH * 4, st) != cudaSuccess || cudaMemsetAsync(p->s2, 0, n * PH * 4, st) != cudaSuccess ||
        cudaMemsetAsync(p->deg, 0, n * 4, st) != cudaSuccess || cudaMemsetAsync(p->mnk, 0xff, n * PH * 8, st) != cudaSuccess ||
        cudaMemsetAsync(p->mxk, 0, n * PH * 8, st) != cudaSuccess)
        return rl_fail(RL_ERR_CUDA, "rl_pna_item_stats: memset", cudaGetLastError());
    const dim3 grid(PNA_BLOCKS, s->num_slots);
    if (fr->count_bits == 32) k_pna_item_stats<uint32_t><<<grid, 256, 0, st>>>(*g, *r, *s, *fr, *c, emb, *p);
    else k_pna_item_stats<unsigned long long><<<grid, 256, 0, st>>>(*g, *r, *s, *fr, *c, emb, *p);
    CHECK_LAUNCH("k_pna_item_stats");
    if (r->num_zero_rules > 0) {
        k_pna_zr_stats<<<s->num_slots, 128, 0, st>>>(*g, *r, *s, *c, emb, *p);
        CHECK_LAUNCH("k_pna_zr_stats");
    }
    return RL_OK;
}

/* qscr: 2 * S * 32 floats of scratch; Y[cap][16] = Linear(12H,H) output (before LayerNorm), FEAT[cap][64], SC[cap] */
int rl_pna_front_forward(const rl_slots *s, const rl_cells *c, const rl_pna *p, const float *W, const float *b, float *qscr,
                         float *Y, float *FEAT, float *SC, void *stream)
{
    if (!s || bad_pna(c, p) || !W || !b || !qscr || !Y || !FEAT || !SC) return rl_fail(RL_ERR_ARG, "rl_pna_front_forward: bad argument");
    if (s->num_slots <= 0) return RL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nq = (size_t)s->num_slots * RL_LANES;
    if (cudaMemsetAsync(qscr, 0, 2 * nq * sizeof(float), st) != cudaSuccess) return rl_fail(RL_ERR_CUDA, "rl_pna_front_forward: memset", cudaGetLastError());
    k_pna_qscale<<<cell_grid(c), 256, 0, st>>>(*c, *p, qscr, qscr + nq);
    CHECK_LAUNCH("k_pna_qscale");
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_pna_front_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_END); attr = true; }
    k_pna_front_fwd<<<tile_grid(c), PT, PF_END, st>>>(*c, *p, qscr, qscr + nq, W, b, Y, FEAT, SC);
    CHECK_LAUNCH("k_pna_front_fwd");
    return RL_OK;
}

/* dstat[cap][64] scratch; gW[16][192] is ACCUMULATED (the bias gradient comes from rl_tail_backward's gb0) */
int rl_pna_front_backward(const rl_cells *c, const rl_pna *p, const float *W, const float *dY, const float *FEAT,
                          const float *SC, float *dstat, float *gW, void *stream)
{
    if (bad_pna(c, p) || !W || !dY || !FEAT || !SC || !dstat || !gW) return rl_fail(RL_ERR_ARG, "rl_pna_front_backward: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_pna_front_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_END); attr = true; }
    k_pna_front_bwd<<<tile_grid(c), PT, PB_END, st>>>(*c, *p, W, dY, FEAT, SC, dstat);
    CHECK_LAUNCH("k_pna_front_bwd");
    static bool attr_w = false;
    if (!attr_w) { cudaFuncSetAttribute(k_pna_w_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_END); attr_w = true; }
    k_pna_w_grad<<<148, PT, PW_END, st>>>(*c, dY, FEAT, SC, gW);
    CHECK_LAUNCH("k_pna_w_grad");
    return RL_OK;
}

int rl_pna_item_backward(const rl_graph *g, const rl_rules *r, const rl_slots *s, const rl_frontier *fr, const rl_cells *c,
                         const float *emb, const rl_pna *p, const float *dstat, float *grad_emb, void *stream)
{
    if (!g || !r || !s || !emb || !dstat || !grad_emb || bad_pna(c, p) || bad_items(fr)) return rl_fail(RL_ERR_ARG, "rl_pna_item_backward: bad argument");
    if (s->num_slots <= 0) return RL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(PNA_BLOCKS, s->num_slots);
    if (fr->count_bits == 32) k_pna_item_bwd<uint32_t><<<grid, 256, 0, st>>>(*g, *r, *s, *fr, *c, emb, *p, dstat, grad_emb);
    else k_pna_item_bwd<unsigned long long><<<grid, 256, 0, st>>>(*g, *r, *s, *fr, *c, emb, *p, dstat, grad_emb);
    CHECK_LAUNCH("k_pna_item_bwd");
    if (r->num_zero_rules > 0) {
        k_pna_zr_bwd<<<s->num_slots, 128, 0, st>>>(*g, *r, *s, *c, emb, *p, dstat, grad_emb);
        CHECK_LAUNCH("k_pna_zr_bwd");
    }
    return RL_OK;
}

}  // extern "C"
