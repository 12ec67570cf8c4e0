// rnnlogic_b200 -- dense tail of PredictorPlus (`sum` aggregator) on the candidate cells, forward and a
// hand-written backward with every weight gradient reduced in-kernel (no [C,128] activation, no library GEMM).
// Reference: src/layers.py:73-75 (Linear(H,H) -> LayerNorm -> ReLU) + src/predictors.py:253-255
// ([.., relation_emb[q]] -> Linear(2H,128) -> ReLU -> Linear(128,1)), H = 16.
//
//   y = W0 F + b0 ; o = relu(LN(y)) ; u = [o, rel[q]] ; a = W1 u + b1 ; z = W2 . relu(a) + b2
//
// Mapping: a warp walks tiles of 16 cells.  In the 2H->128 layer every LANE OWNS FOUR HIDDEN UNITS and keeps
// their 4 x 32 weights in registers; a cell's input vector u is broadcast from shared memory (8 LDS.128 for
// 128 FMA per lane -- a thread-per-cell mapping reads one shared-memory weight per FMA and is bound by the
// LDS pipe).  The small H x H front runs on (cell, unit) pairs, two cells per warp pass, LayerNorm statistics
// by shuffles inside the half warp.  fp32 FFMA throughout: the parity bar is 1e-5 on the logits, TF32
// tensor-core MMA (10-bit mantissa) is out, and at K = 32 a 3xTF32 split buys nothing over FFMA.
//
// Backward per tile: phase A recomputes a, forms d1 = relu'(a) * g * W2 (kept in shared memory), reduces
// du[0..16) across the lanes with a 16-shuffle transpose-reduction and accumulates sum_cells d1 per head
// relation (the relation-embedding half of du is linear in it); phase B re-uses the registers of the W1 tile
// as 4 x 32 accumulators of dW1 += d1 (x) u and flushes them to a block accumulator in shared memory; the
// front backward (LayerNorm, Linear(H,H)) runs on (cell, unit) pairs again.  One atomicAdd per weight
// gradient element and block at the end.
#include "rl_device.cuh"

#define TH 16         // hidden_dim
#define TJ 128        // hidden width of the score MLP
#define TK 32         // its input width (2H)
#define TT 16         // cells per warp tile
#define TW 8          // warps per block

struct TailP { const float *W0, *b0, *gamma, *beta, *W1, *b1, *W2, *b2, *rel; };
struct TailG { float *W0, *b0, *gamma, *beta, *W1, *b1, *W2, *b2, *d1sum; };

// shared-memory layout (floats)
#define SM_W1T 0                          // [TK][TJ]  W1 transposed: lane reads its 4 units of input k with one LDS.128
#define SM_B1 (SM_W1T + TK * TJ)
#define SM_W2 (SM_B1 + TJ)
#define SM_W0 (SM_W2 + TJ)                // [TH][TH+1] padded rows
#define SM_B0 (SM_W0 + TH * (TH + 1))
#define SM_GA (SM_B0 + TH)
#define SM_BE (SM_GA + TH)
#define SM_FWD_END (SM_BE + TH)
#define SM_DW1 SM_FWD_END                 // backward: [(k*4+jj)*32 + lane] block accumulator of dW1
#define SM_SMALL (SM_DW1 + TK * TJ)       // dW2[TJ] db1[TJ] dW0[TH*TH] db0[TH] dgamma[TH] dbeta[TH] db2[1]
#define SM_SMALL_N (2 * TJ + TH * TH + 3 * TH + 1)
#define SM_BWD_END (SM_SMALL + ((SM_SMALL_N + 3) & ~3))
#define WARP_FWD (TT * TK)                                    // U
#define WARP_BWD (TT * TK + TT * TJ + TT * TH + TT * TH + 3 * TT)   // U, D, DU, NRM, RSTD, GV, HEAD

__device__ __forceinline__ float half_sum(float v)            // sum over the 16 lanes of a half warp
{
    v += __shfl_xor_sync(FULL, v, 8);
    v += __shfl_xor_sync(FULL, v, 4);
    v += __shfl_xor_sync(FULL, v, 2);
    v += __shfl_xor_sync(FULL, v, 1);
    return v;
}

// one step of the transpose-reduction: lanes whose BIT is set keep the upper HALF of the vector, the others the
// lower one; each sends the half it drops to its partner (lane ^ BIT)
template <int HALF, int BIT>
__device__ __forceinline__ void reduce_step(float (&v)[TH], int lane)
{
    const bool up = (lane & BIT) != 0;
#pragma unroll
    for (int m = 0; m < HALF; ++m) {
        const float lo = v[m], hi = v[m + HALF];
        const float recv = __shfl_xor_sync(FULL, up ? lo : hi, BIT);
        v[m] = (up ? hi : lo) + recv;
    }
}

__device__ __forceinline__ void stage_weights(float *sm, const TailP &w, int tid, int nthreads)
{
    for (int i = tid; i < TJ * TK; i += nthreads) {           // W1 [j][k] row-major -> [k][j]
        const int j = i / TK, k = i % TK;
        sm[SM_W1T + k * TJ + j] = w.W1[i];
    }
    for (int i = tid; i < TJ; i += nthreads) { sm[SM_B1 + i] = w.b1[i]; sm[SM_W2 + i] = w.W2[i]; }
    for (int i = tid; i < TH * TH; i += nthreads) sm[SM_W0 + (i / TH) * (TH + 1) + (i % TH)] = w.W0[i];
    for (int i = tid; i < TH; i += nthreads) { sm[SM_B0 + i] = w.b0[i]; sm[SM_GA + i] = w.gamma[i]; sm[SM_BE + i] = w.beta[i]; }
}

// Front of one tile: (cell, unit) pairs, two cells per pass.  Writes u = [relu(LN(W0 F + b0)), rel[head]] of the
// tile's cells to U; the backward also keeps the normalised values, 1/std and the head relation.
template <bool BWD>
__device__ __forceinline__ void tile_front(const float *sm, const float *__restrict__ F, const int32_t *__restrict__ cell_key,
                                           const int32_t *__restrict__ slot_head, const float *__restrict__ rel, long long cell0,
                                           long long C, float *U, float *NRM, float *RSTD, int *HEAD)
{
    const int lane = threadIdx.x & 31;
    const int i = lane & 15;
#pragma unroll 2
    for (int it = 0; it < TT / 2; ++it) {
        const int cl = it * 2 + (lane >> 4);
        const long long cell = cell0 + cl;
        const bool ok = cell < C;
        float f[TH];
        const float4 *fp = reinterpret_cast<const float4 *>(F + (ok ? cell : 0) * TH);
#pragma unroll
        for (int k4 = 0; k4 < TH / 4; ++k4) {
            const float4 v = ok ? __ldg(fp + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
            f[4 * k4] = v.x; f[4 * k4 + 1] = v.y; f[4 * k4 + 2] = v.z; f[4 * k4 + 3] = v.w;
        }
        float y = sm[SM_B0 + i];
#pragma unroll
        for (int k = 0; k < TH; ++k) y = fmaf(sm[SM_W0 + i * (TH + 1) + k], f[k], y);
        const float mean = half_sum(y) / (float)TH;
        const float d = y - mean;
        const float var = half_sum(d * d) / (float)TH;
        const float rstd = rsqrtf(var + 1e-5f);
        const float nrm = d * rstd;
        const float o = fmaxf(fmaf(sm[SM_GA + i], nrm, sm[SM_BE + i]), 0.f);
        const int head = ok ? slot_head[cell_key[cell] >> 5] : -1;
        U[cl * TK + i] = ok ? o : 0.f;
        U[cl * TK + TH + i] = ok ? __ldg(rel + (size_t)head * TH + i) : 0.f;
        if (BWD) {
            NRM[cl * TH + i] = nrm;
            if (i == 0) { RSTD[cl] = rstd; HEAD[cl] = head; }
        }
    }
    __syncwarp();
}

// a[jj] = b1 + sum_k W1[4*lane+jj][k] * u[k] for the lane's four hidden units
__device__ __forceinline__ float4 hidden_pre(const float4 (&w1v)[TK], const float4 b1v, const float *Uc)
{
    float4 a = b1v;
    const float4 *u4 = reinterpret_cast<const float4 *>(Uc);
#pragma unroll
    for (int k4 = 0; k4 < TK / 4; ++k4) {
        const float4 u = u4[k4];
        a.x = fmaf(w1v[4 * k4].x, u.x, a.x); a.y = fmaf(w1v[4 * k4].y, u.x, a.y); a.z = fmaf(w1v[4 * k4].z, u.x, a.z); a.w = fmaf(w1v[4 * k4].w, u.x, a.w);
        a.x = fmaf(w1v[4 * k4 + 1].x, u.y, a.x); a.y = fmaf(w1v[4 * k4 + 1].y, u.y, a.y); a.z = fmaf(w1v[4 * k4 + 1].z, u.y, a.z); a.w = fmaf(w1v[4 * k4 + 1].w, u.y, a.w);
        a.x = fmaf(w1v[4 * k4 + 2].x, u.z, a.x); a.y = fmaf(w1v[4 * k4 + 2].y, u.z, a.y); a.z = fmaf(w1v[4 * k4 + 2].z, u.z, a.z); a.w = fmaf(w1v[4 * k4 + 2].w, u.z, a.w);
        a.x = fmaf(w1v[4 * k4 + 3].x, u.w, a.x); a.y = fmaf(w1v[4 * k4 + 3].y, u.w, a.y); a.z = fmaf(w1v[4 * k4 + 3].z, u.w, a.z); a.w = fmaf(w1v[4 * k4 + 3].w, u.w, a.w);
    }
    return a;
}

__global__ void __launch_bounds__(TW * 32, 1)
k_tail_fwd(const int32_t *__restrict__ counters, int cap, const float *__restrict__ F, const int32_t *__restrict__ cell_key,
           const int32_t *__restrict__ slot_head, TailP w, float *__restrict__ zc)
{
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long C = min(counters[0], cap);
    if ((long long)blockIdx.x * TW * TT >= C) return;
    stage_weights(sm, w, threadIdx.x, TW * 32);
    __syncthreads();
    float4 w1v[TK];
#pragma unroll
    for (int k = 0; k < TK; ++k) w1v[k] = *reinterpret_cast<const float4 *>(sm + SM_W1T + k * TJ + 4 * lane);
    const float4 b1v = *reinterpret_cast<const float4 *>(sm + SM_B1 + 4 * lane);
    const float4 w2v = *reinterpret_cast<const float4 *>(sm + SM_W2 + 4 * lane);
    const float b2 = __ldg(w.b2);
    float *U = sm + SM_FWD_END + warp * WARP_FWD;
    for (long long tile = (long long)blockIdx.x * TW + warp; tile * TT < C; tile += (long long)gridDim.x * TW) {
        const long long cell0 = tile * TT;
        tile_front<false>(sm, F, cell_key, slot_head, w.rel, cell0, C, U, nullptr, nullptr, nullptr);
        float zmine = 0.f;                                    // lane c keeps the score of cell c of the tile
#pragma unroll 2
        for (int c = 0; c < TT; ++c) {
            const float4 a = hidden_pre(w1v, b1v, U + c * TK);
            float zp = w2v.x * fmaxf(a.x, 0.f);
            zp = fmaf(w2v.y, fmaxf(a.y, 0.f), zp);
            zp = fmaf(w2v.z, fmaxf(a.z, 0.f), zp);
            zp = fmaf(w2v.w, fmaxf(a.w, 0.f), zp);
            zp = warp_sumf(zp);
            if (lane == c) zmine = zp + b2;
        }
        if (lane < TT && cell0 + lane < C) zc[cell0 + lane] = zmine;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(TW * 32, 1)
k_tail_bwd(const int32_t *__restrict__ counters, int cap, const float *__restrict__ F, const int32_t *__restrict__ cell_key,
           const int32_t *__restrict__ slot_head, TailP w, const float *__restrict__ Gc, float *__restrict__ dF, TailG gr)
{
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long C = min(counters[0], cap);
    if ((long long)blockIdx.x * TW * TT >= C) return;
    stage_weights(sm, w, threadIdx.x, TW * 32);
    for (int i = threadIdx.x; i < TK * TJ + ((SM_SMALL_N + 3) & ~3); i += TW * 32) sm[SM_DW1 + i] = 0.f;
    __syncthreads();
    const float4 b1v = *reinterpret_cast<const float4 *>(sm + SM_B1 + 4 * lane);
    const float4 w2v = *reinterpret_cast<const float4 *>(sm + SM_W2 + 4 * lane);
    float *U = sm + SM_BWD_END + warp * WARP_BWD;
    float *D = U + TT * TK, *DU = D + TT * TJ, *NRM = DU + TT * TH, *RSTD = NRM + TT * TH, *GV = RSTD + TT;
    int *HEAD = reinterpret_cast<int *>(GV + TT);
    // lane-private accumulators that live for the whole kernel
    float4 dW2a = make_float4(0.f, 0.f, 0.f, 0.f), db1a = dW2a, d1h = dW2a;
    float dW0a[TH];
#pragma unroll
    for (int k = 0; k < TH; ++k) dW0a[k] = 0.f;
    float dga = 0.f, dbta = 0.f, db0a = 0.f, db2a = 0.f;
    int cur_head = -1;
    const int i16 = lane & 15;
    auto flush_head = [&]() {
        if (cur_head >= 0) {
            float *p = gr.d1sum + (size_t)cur_head * TJ + 4 * lane;
            if (d1h.x != 0.f) atomicAdd(p, d1h.x);
            if (d1h.y != 0.f) atomicAdd(p + 1, d1h.y);
            if (d1h.z != 0.f) atomicAdd(p + 2, d1h.z);
            if (d1h.w != 0.f) atomicAdd(p + 3, d1h.w);
        }
        d1h = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    for (long long tile = (long long)blockIdx.x * TW + warp; tile * TT < C; tile += (long long)gridDim.x * TW) {
        const long long cell0 = tile * TT;
        if (lane < TT) GV[lane] = cell0 + lane < C ? Gc[cell0 + lane] : 0.f;
        tile_front<true>(sm, F, cell_key, slot_head, w.rel, cell0, C, U, NRM, RSTD, HEAD);
        // ---- phase A: hidden pre-activations, d1, du[0..16) -------------------------------------------
        {
            float4 w1v[TK];
#pragma unroll
            for (int k = 0; k < TK; ++k) w1v[k] = *reinterpret_cast<const float4 *>(sm + SM_W1T + k * TJ + 4 * lane);
            for (int c = 0; c < TT; ++c) {
                const float4 a = hidden_pre(w1v, b1v, U + c * TK);
                const float gq = GV[c];
                const int head = HEAD[c];
                if (head != cur_head) {                          // warp-uniform: cells are ordered by slot
                    flush_head();
                    cur_head = head;
                }
                float4 d1;
                d1.x = a.x > 0.f ? gq * w2v.x : 0.f;
                d1.y = a.y > 0.f ? gq * w2v.y : 0.f;
                d1.z = a.z > 0.f ? gq * w2v.z : 0.f;
                d1.w = a.w > 0.f ? gq * w2v.w : 0.f;
                dW2a.x = fmaf(gq, fmaxf(a.x, 0.f), dW2a.x); dW2a.y = fmaf(gq, fmaxf(a.y, 0.f), dW2a.y);
                dW2a.z = fmaf(gq, fmaxf(a.z, 0.f), dW2a.z); dW2a.w = fmaf(gq, fmaxf(a.w, 0.f), dW2a.w);
                db1a.x += d1.x; db1a.y += d1.y; db1a.z += d1.z; db1a.w += d1.w;
                d1h.x += d1.x; d1h.y += d1.y; d1h.z += d1.z; d1h.w += d1.w;
                db2a += gq;
                *reinterpret_cast<float4 *>(D + c * TJ + 4 * lane) = d1;
                // du[k] = sum_j d1[j] W1[j][k], k < 16: the lane's share, then a transpose-reduction over the
                // lanes (16 shuffles instead of 16 x 5); lane L ends with du[L >> 1]
                float v[TH];
#pragma unroll
                for (int k = 0; k < TH; ++k)
                    v[k] = fmaf(d1.w, w1v[k].w, fmaf(d1.z, w1v[k].z, fmaf(d1.y, w1v[k].y, d1.x * w1v[k].x)));
                reduce_step<8, 16>(v, lane);
                reduce_step<4, 8>(v, lane);
                reduce_step<2, 4>(v, lane);
                reduce_step<1, 2>(v, lane);
                v[0] += __shfl_xor_sync(FULL, v[0], 1);
                if (!(lane & 1)) DU[c * TH + (lane >> 1)] = v[0];
            }
        }
        __syncwarp();
        // ---- phase B: dW1 += d1 (x) u, accumulators in the registers the W1 tile used -----------------
        {
            float4 acc[TK];
#pragma unroll
            for (int k = 0; k < TK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = 0; c < TT; ++c) {
                const float4 d1 = *reinterpret_cast<const float4 *>(D + c * TJ + 4 * lane);
                const float4 *u4 = reinterpret_cast<const float4 *>(U + c * TK);
#pragma unroll
                for (int k4 = 0; k4 < TK / 4; ++k4) {
                    const float4 u = u4[k4];
                    acc[4 * k4].x = fmaf(d1.x, u.x, acc[4 * k4].x); acc[4 * k4].y = fmaf(d1.y, u.x, acc[4 * k4].y);
                    acc[4 * k4].z = fmaf(d1.z, u.x, acc[4 * k4].z); acc[4 * k4].w = fmaf(d1.w, u.x, acc[4 * k4].w);
                    acc[4 * k4 + 1].x = fmaf(d1.x, u.y, acc[4 * k4 + 1].x); acc[4 * k4 + 1].y = fmaf(d1.y, u.y, acc[4 * k4 + 1].y);
                    acc[4 * k4 + 1].z = fmaf(d1.z, u.y, acc[4 * k4 + 1].z); acc[4 * k4 + 1].w = fmaf(d1.w, u.y, acc[4 * k4 + 1].w);
                    acc[4 * k4 + 2].x = fmaf(d1.x, u.z, acc[4 * k4 + 2].x); acc[4 * k4 + 2].y = fmaf(d1.y, u.z, acc[4 * k4 + 2].y);
                    acc[4 * k4 + 2].z = fmaf(d1.z, u.z, acc[4 * k4 + 2].z); acc[4 * k4 + 2].w = fmaf(d1.w, u.z, acc[4 * k4 + 2].w);
                    acc[4 * k4 + 3].x = fmaf(d1.x, u.w, acc[4 * k4 + 3].x); acc[4 * k4 + 3].y = fmaf(d1.y, u.w, acc[4 * k4 + 3].y);
                    acc[4 * k4 + 3].z = fmaf(d1.z, u.w, acc[4 * k4 + 3].z); acc[4 * k4 + 3].w = fmaf(d1.w, u.w, acc[4 * k4 + 3].w);
                }
            }
            float *dst = sm + SM_DW1 + lane;
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                atomicAdd(dst + (k * 4 + 0) * 32, acc[k].x);
                atomicAdd(dst + (k * 4 + 1) * 32, acc[k].y);
                atomicAdd(dst + (k * 4 + 2) * 32, acc[k].z);
                atomicAdd(dst + (k * 4 + 3) * 32, acc[k].w);
            }
        }
        // ---- front backward: ReLU, LayerNorm, Linear(H,H) on (cell, unit) pairs ----------------------
#pragma unroll 2
        for (int it = 0; it < TT / 2; ++it) {
            const int cl = it * 2 + (lane >> 4);
            const long long cell = cell0 + cl;
            const bool ok = cell < C;
            const float nrm = NRM[cl * TH + i16], rstd = RSTD[cl];
            const float ga = sm[SM_GA + i16];
            const float pre = fmaf(ga, nrm, sm[SM_BE + i16]);
            const float d_o = pre > 0.f ? DU[cl * TH + i16] : 0.f;
            dga = fmaf(d_o, nrm, dga);
            dbta += d_o;
            const float dn = d_o * ga;
            const float m1 = half_sum(dn) / (float)TH;
            const float m2 = half_sum(dn * nrm) / (float)TH;
            const float dy = rstd * (dn - m1 - nrm * m2);
            db0a += dy;
            const float4 *fp = reinterpret_cast<const float4 *>(F + (ok ? cell : 0) * TH);
#pragma unroll
            for (int k4 = 0; k4 < TH / 4; ++k4) {
                const float4 fv = ok ? __ldg(fp + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
                dW0a[4 * k4] = fmaf(dy, fv.x, dW0a[4 * k4]); dW0a[4 * k4 + 1] = fmaf(dy, fv.y, dW0a[4 * k4 + 1]);
                dW0a[4 * k4 + 2] = fmaf(dy, fv.z, dW0a[4 * k4 + 2]); dW0a[4 * k4 + 3] = fmaf(dy, fv.w, dW0a[4 * k4 + 3]);
            }
            float dx = 0.f;                                     // dF[k = i16] = sum_i W0[i][k] dy_i
#pragma unroll
            for (int ii = 0; ii < TH; ++ii)
                dx = fmaf(sm[SM_W0 + ii * (TH + 1) + i16], __shfl_sync(FULL, dy, (lane & 16) + ii), dx);
            if (ok) dF[cell * TH + i16] = dx;
        }
        __syncwarp();
    }
    flush_head();
    // ---- block reduction of the lane-private accumulators, then one atomic per element -----------------
    float *sml = sm + SM_SMALL;
    atomicAdd(sml + 4 * lane, dW2a.x); atomicAdd(sml + 4 * lane + 1, dW2a.y);
    atomicAdd(sml + 4 * lane + 2, dW2a.z); atomicAdd(sml + 4 * lane + 3, dW2a.w);
    atomicAdd(sml + TJ + 4 * lane, db1a.x); atomicAdd(sml + TJ + 4 * lane + 1, db1a.y);
    atomicAdd(sml + TJ + 4 * lane + 2, db1a.z); atomicAdd(sml + TJ + 4 * lane + 3, db1a.w);
#pragma unroll
    for (int k = 0; k < TH; ++k) atomicAdd(sml + 2 * TJ + i16 * TH + k, dW0a[k]);
    atomicAdd(sml + 2 * TJ + TH * TH + i16, db0a);
    atomicAdd(sml + 2 * TJ + TH * TH + TH + i16, dga);
    atomicAdd(sml + 2 * TJ + TH * TH + 2 * TH + i16, dbta);
    if (lane == 0) atomicAdd(sml + 2 * TJ + TH * TH + 3 * TH, db2a);
    __syncthreads();
    for (int i = threadIdx.x; i < TK * TJ; i += TW * 32) {       // sm index (k*4+jj)*32 + l  ->  W1[4*l+jj][k]
        const int l = i & 31, jj = (i >> 5) & 3, k = i >> 7;
        const float v = sm[SM_DW1 + i];
        if (v != 0.f) atomicAdd(gr.W1 + (4 * l + jj) * TK + k, v);
    }
    for (int i = threadIdx.x; i < SM_SMALL_N; i += TW * 32) {
        const float v = sml[i];
        if (v == 0.f) continue;
        float *dst;
        if (i < TJ) dst = gr.W2 + i;
        else if (i < 2 * TJ) dst = gr.b1 + (i - TJ);
        else if (i < 2 * TJ + TH * TH) dst = gr.W0 + (i - 2 * TJ);
        else if (i < 2 * TJ + TH * TH + TH) dst = gr.b0 + (i - 2 * TJ - TH * TH);
        else if (i < 2 * TJ + TH * TH + 2 * TH) dst = gr.gamma + (i - 2 * TJ - TH * TH - TH);
        else if (i < 2 * TJ + TH * TH + 3 * TH) dst = gr.beta + (i - 2 * TJ - TH * TH - 2 * TH);
        else dst = gr.b2;
        atomicAdd(dst, v);
    }
}

// relation-embedding gradient: du[16+k] = sum_j d1[j] W1[j][16+k] is linear in d1, so the sum over the cells of a
// head relation only needs sum_cells d1 (d1sum[head][j], accumulated by k_tail_bwd)
__global__ void __launch_bounds__(TH)
k_tail_rel_grad(const float *__restrict__ W1, const float *__restrict__ d1sum, float *__restrict__ grad_rel)
{
    const int head = blockIdx.x, k = threadIdx.x;
    const float *d = d1sum + (size_t)head * TJ;
    float acc = 0.f;
    for (int j = 0; j < TJ; ++j) acc = fmaf(d[j], W1[j * TK + TH + k], acc);
    if (acc != 0.f) grad_rel[(size_t)head * TH + k] += acc;
}

static int tail_blocks()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

extern "C" {

int rl_tail_forward(const rl_cells *c, const int32_t *slot_head, int32_t H, int32_t J, const float *F, const float *W0,
                    const float *b0, const float *gamma, const float *beta, const float *W1, const float *b1,
                    const float *W2, const float *b2, const float *rel_emb, float *zc, void *stream)
{
    if (!c || !c->counters || !c->cell_key || !slot_head || !F || !W0 || !b0 || !gamma || !beta || !W1 || !b1 || !W2 || !b2 || !rel_emb || !zc)
        return rl_fail(RL_ERR_ARG, "rl_tail_forward: null argument");
    if (H != TH || J != TJ) return rl_fail(RL_ERR_ARG, "rl_tail_forward: built for hidden_dim 16 and a 128-wide score MLP");
    TailP w{W0, b0, gamma, beta, W1, b1, W2, b2, rel_emb};
    const size_t smem = (size_t)(SM_FWD_END + TW * WARP_FWD) * sizeof(float);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_tail_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    k_tail_fwd<<<tail_blocks(), TW * 32, smem, (cudaStream_t)stream>>>(c->counters, c->cap, F, c->cell_key, slot_head, w, zc);
    CHECK_LAUNCH("k_tail_fwd");
    return RL_OK;
}

/* grads are ACCUMULATED (atomicAdd); d1sum [R][128] is scratch */
int rl_tail_backward(const rl_cells *c, const int32_t *slot_head, int32_t R, int32_t H, int32_t J, const float *F,
                     const float *W0, const float *b0, const float *gamma, const float *beta, const float *W1,
                     const float *b1, const float *W2, const float *b2, const float *rel_emb, const float *Gc,
                     float *dF, float *gW0, float *gb0, float *ggamma, float *gbeta, float *gW1, float *gb1,
                     float *gW2, float *gb2, float *grel, float *d1sum, void *stream)
{
    if (!c || !c->counters || !c->cell_key || !slot_head || !F || !W0 || !b0 || !gamma || !beta || !W1 || !b1 || !W2 || !b2 || !rel_emb ||
        !Gc || !dF || !gW0 || !gb0 || !ggamma || !gbeta || !gW1 || !gb1 || !gW2 || !gb2 || !grel || !d1sum || R <= 0)
        return rl_fail(RL_ERR_ARG, "rl_tail_backward: null argument");
    if (H != TH || J != TJ) return rl_fail(RL_ERR_ARG, "rl_tail_backward: built for hidden_dim 16 and a 128-wide score MLP");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d1sum, 0, (size_t)R * TJ * sizeof(float), st);
    if (e != cudaSuccess) return rl_fail(RL_ERR_CUDA, "rl_tail_backward: memset", e);
    TailP w{W0, b0, gamma, beta, W1, b1, W2, b2, rel_emb};
    TailG gr{gW0, gb0, ggamma, gbeta, gW1, gb1, gW2, gb2, d1sum};
    const size_t smem = (size_t)(SM_BWD_END + TW * WARP_BWD) * sizeof(float);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_tail_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    k_tail_bwd<<<tail_blocks(), TW * 32, smem, st>>>(c->counters, c->cap, F, c->cell_key, slot_head, w, Gc, dF, gr);
    CHECK_LAUNCH("k_tail_bwd");
    k_tail_rel_grad<<<R, TH, 0, st>>>(W1, d1sum, grel);
    CHECK_LAUNCH("k_tail_rel_grad");
    return RL_OK;
}

}  // extern "C"
