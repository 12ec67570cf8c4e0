// rnnlogic_b200 -- dense tail of PredictorPlus (`sum` aggregator) on the candidate cells, forward and a
// hand-written backward with every weight gradient reduced in-kernel (no [C,128] activation, no library GEMM).
// Reference: src/layers.py:73-75 (Linear(H,H) -> LayerNorm -> ReLU) + src/predictors.py:253-255
// ([.., relation_emb[q]] -> Linear(2H,128) -> ReLU -> Linear(128,1)), H = 16, J = 128.
//
//   y = W0 F + b0 ; o = relu(LN(y)) ; u = [o, rel[q]] ; a = W1 u + b1 ; z = W2 . relu(a) + b2
//
// fp32 FFMA throughout: the parity bar is 1e-5 on the logits, TF32 tensor-core MMA (10-bit mantissa) is out and
// at K = 32 a 3xTF32 split buys nothing over FFMA.  Three kernels:
//
//  k_tailc_fwd        one THREAD per cell (weights broadcast from shared memory with LDS.128, two hidden units in
//                     flight): z, the front output o[16] and the 128 ReLU bits of the hidden layer (16 bytes).
//  k_tailc_bwd_cells  one thread per cell: with the ReLU bits the hidden layer needs no recomputation --
//                     du[k] = g * sum_j bit_j * (W2_j W1[j][k]) for the 16 inputs that continue into the front
//                     (the relation half of du is only ever summed per head relation, see below) -- then the
//                     LayerNorm / Linear(H,H) backward: dF[16], dy[16], and the small bias / LayerNorm gradients.
//  k_tailc_bwd_weights every LANE OWNS FOUR HIDDEN UNITS and keeps V[4][32] = sum_cells bit_j * g * u_k and
//                     P[4] = sum_cells bit_j * g in registers for the whole kernel (no flush, no atomics inside
//                     the loop; a cell's g*u is broadcast from shared memory: 8 LDS.128 per 128 FMA).  Everything
//                     else follows from V and P:
//                         dW1[j][k] = W2_j V[j][k]       db1[j] = W2_j P[j]
//                         dW2[j]    = sum_cells g relu(a_j) = b1_j P[j] + sum_k W1[j][k] V[j][k]
//                         drel[q][k] = sum_j W1[j][16+k] W2_j P_q[j]      (P per head relation q)
//                     and dW0 = dy^T F rides along on (unit, half) lane pairs.
#include "rl_device.cuh"

#define TH 16         // hidden_dim
#define TJ 128        // hidden width of the score MLP
#define TK 32         // its input width (2H)
#define TT 16         // cells per warp tile of k_tailc_bwd_weights
#define TW 8          // warps per block of k_tailc_bwd_weights

struct TailP { const float *W0, *b0, *gamma, *beta, *W1, *b1, *W2, *b2, *rel; };

// ---- shared-memory layout of the per-cell kernels (floats) ----
#define SC_W1 0                           // forward: W1 [TJ][TK];  backward: W21 [TJ][TH] = W2_j * W1[j][k<16]
#define SC_B1 (SC_W1 + TJ * TK)
#define SC_W2 (SC_B1 + TJ)
#define SC_W0 (SC_W2 + TJ)                // [TH][TH]
#define SC_B0 (SC_W0 + TH * TH)
#define SC_GA (SC_B0 + TH)
#define SC_BE (SC_GA + TH)
#define SC_ACC (SC_BE + TH)               // backward: db0[TH] dgamma[TH] dbeta[TH] db2[1]
#define SC_END (SC_ACC + 3 * TH + 4)

// front of one cell in registers: y = W0 f + b0, LayerNorm statistics; returns 1/std, fills nrm[]
template <bool PRE>
__device__ __forceinline__ float cell_front(const float *sm, const float (&f)[TH], float (&nrm)[TH])
{
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < TH; ++i) {
        if (PRE) {                                               // the aggregator's Linear was applied by its own front kernel (PNA)
            nrm[i] = f[i];
            mean += f[i];
            continue;
        }
        float a = sm[SC_B0 + i];
        const float4 *wr = reinterpret_cast<const float4 *>(sm + SC_W0 + i * TH);
#pragma unroll
        for (int k4 = 0; k4 < TH / 4; ++k4) {
            const float4 w4 = wr[k4];
            a = fmaf(w4.x, f[4 * k4], a); a = fmaf(w4.y, f[4 * k4 + 1], a);
            a = fmaf(w4.z, f[4 * k4 + 2], a); a = fmaf(w4.w, f[4 * k4 + 3], a);
        }
        nrm[i] = a;
        mean += a;
    }
    mean /= (float)TH;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < TH; ++i) { const float d = nrm[i] - mean; var = fmaf(d, d, var); }
    var /= (float)TH;
    const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll
    for (int i = 0; i < TH; ++i) nrm[i] = (nrm[i] - mean) * rstd;
    return rstd;
}

__device__ __forceinline__ void load_row16(const float *__restrict__ p, float (&f)[TH])
{
    const float4 *fp = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int k4 = 0; k4 < TH / 4; ++k4) {
        const float4 v = __ldg(fp + k4);
        f[4 * k4] = v.x; f[4 * k4 + 1] = v.y; f[4 * k4 + 2] = v.z; f[4 * k4 + 3] = v.w;
    }
}

template <bool PRE>
__global__ void __launch_bounds__(256, 2)
k_tailc_fwd(const int32_t *__restrict__ counters, int cap, const float *__restrict__ F, const int32_t *__restrict__ cell_key,
            const int32_t *__restrict__ slot_head, TailP w, float *__restrict__ zc, float *__restrict__ O,
            uint4 *__restrict__ bits)
{
    extern __shared__ __align__(16) float sm[];
    const long long C = min(counters[0], cap);
    if ((long long)blockIdx.x * 256 >= C) return;
    for (int i = threadIdx.x; i < TJ * TK; i += 256) sm[SC_W1 + i] = w.W1[i];
    for (int i = threadIdx.x; i < TJ; i += 256) { sm[SC_B1 + i] = w.b1[i]; sm[SC_W2 + i] = w.W2[i]; }
    for (int i = threadIdx.x; i < TH * TH; i += 256) sm[SC_W0 + i] = w.W0[i];
    for (int i = threadIdx.x; i < TH; i += 256) { sm[SC_B0 + i] = w.b0[i]; sm[SC_GA + i] = w.gamma[i]; sm[SC_BE + i] = w.beta[i]; }
    __syncthreads();
    const float b2 = __ldg(w.b2);
    for (long long cell = (long long)blockIdx.x * 256 + threadIdx.x; cell < C; cell += (long long)gridDim.x * 256) {
        float f[TH], nrm[TH], u[TK];
        load_row16(F + cell * TH, f);
        cell_front<PRE>(sm, f, nrm);
#pragma unroll
        for (int i = 0; i < TH; ++i) u[i] = fmaxf(fmaf(sm[SC_GA + i], nrm[i], sm[SC_BE + i]), 0.f);
        float4 *op = reinterpret_cast<float4 *>(O + cell * TH);
#pragma unroll
        for (int k4 = 0; k4 < TH / 4; ++k4) op[k4] = make_float4(u[4 * k4], u[4 * k4 + 1], u[4 * k4 + 2], u[4 * k4 + 3]);
        const int head = slot_head[cell_key[cell] >> 5];
        {
            float r16[TH];
            load_row16(w.rel + (size_t)head * TH, r16);
#pragma unroll
            for (int i = 0; i < TH; ++i) u[TH + i] = r16[i];
        }
        float z = b2;
        uint32_t word[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
        for (int jw = 0; jw < 4; ++jw) {
            uint32_t wbits = 0u;
#pragma unroll 4
            for (int jj = 0; jj < 32; jj += 2) {                 // two hidden units in flight
                const int j = jw * 32 + jj;
                float a0 = sm[SC_B1 + j], a1 = sm[SC_B1 + j + 1];
                const float4 *w0 = reinterpret_cast<const float4 *>(sm + SC_W1 + j * TK);
                const float4 *w1 = w0 + TK / 4;
#pragma unroll
                for (int k4 = 0; k4 < TK / 4; ++k4) {
                    const float4 p = w0[k4], q = w1[k4];
                    a0 = fmaf(p.x, u[4 * k4], a0); a1 = fmaf(q.x, u[4 * k4], a1);
                    a0 = fmaf(p.y, u[4 * k4 + 1], a0); a1 = fmaf(q.y, u[4 * k4 + 1], a1);
                    a0 = fmaf(p.z, u[4 * k4 + 2], a0); a1 = fmaf(q.z, u[4 * k4 + 2], a1);
                    a0 = fmaf(p.w, u[4 * k4 + 3], a0); a1 = fmaf(q.w, u[4 * k4 + 3], a1);
                }
                z = fmaf(sm[SC_W2 + j], fmaxf(a0, 0.f), z);
                z = fmaf(sm[SC_W2 + j + 1], fmaxf(a1, 0.f), z);
                wbits |= (a0 > 0.f ? 1u : 0u) << jj;
                wbits |= (a1 > 0.f ? 1u : 0u) << (jj + 1);
            }
            word[jw] = wbits;
        }
        zc[cell] = z;
        bits[cell] = make_uint4(word[0], word[1], word[2], word[3]);
    }
}

template <bool PRE>
__global__ void __launch_bounds__(256, 2)
k_tailc_bwd_cells(const int32_t *__restrict__ counters, int cap, const float *__restrict__ F, TailP w,
                  const float *__restrict__ Gc, const uint4 *__restrict__ bits, float *__restrict__ dF,
                  float *__restrict__ dY, float *__restrict__ gb0, float *__restrict__ ggamma, float *__restrict__ gbeta,
                  float *__restrict__ gb2)
{
    extern __shared__ __align__(16) float sm[];
    const long long C = min(counters[0], cap);
    if ((long long)blockIdx.x * 256 >= C) return;
    for (int i = threadIdx.x; i < TJ * TH; i += 256) {           // W21[j][k] = W2_j * W1[j][k], k < 16
        const int j = i / TH, k = i % TH;
        sm[SC_W1 + i] = w.W2[j] * w.W1[j * TK + k];
    }
    for (int i = threadIdx.x; i < TH * TH; i += 256) sm[SC_W0 + i] = w.W0[i];
    for (int i = threadIdx.x; i < TH; i += 256) { sm[SC_B0 + i] = w.b0[i]; sm[SC_GA + i] = w.gamma[i]; sm[SC_BE + i] = w.beta[i]; }
    for (int i = threadIdx.x; i < 3 * TH + 4; i += 256) sm[SC_ACC + i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long span = (long long)gridDim.x * 256;
    for (long long base = (long long)blockIdx.x * 256; base < C; base += span) {    // whole warps stay in the loop (shuffles)
        const long long cell = base + threadIdx.x;
        const bool live = cell < C;
        float f[TH], nrm[TH];
        if (live) load_row16(F + cell * TH, f);
        else {
#pragma unroll
            for (int k = 0; k < TH; ++k) f[k] = 0.f;
        }
        const float rstd = cell_front<PRE>(sm, f, nrm);
        const float gq = live ? Gc[cell] : 0.f;
        const uint4 bw = live ? bits[cell] : make_uint4(0u, 0u, 0u, 0u);
        const uint32_t word[4] = {bw.x, bw.y, bw.z, bw.w};
        float du[TH];
#pragma unroll
        for (int k = 0; k < TH; ++k) du[k] = 0.f;
#pragma unroll 1
        for (int jw = 0; jw < 4; ++jw) {
            const uint32_t wb = word[jw];
#pragma unroll 8
            for (int jj = 0; jj < 32; ++jj) {
                const float bf = (float)((wb >> jj) & 1u);
                const float4 *wr = reinterpret_cast<const float4 *>(sm + SC_W1 + (jw * 32 + jj) * TH);
#pragma unroll
                for (int k4 = 0; k4 < TH / 4; ++k4) {
                    const float4 w4 = wr[k4];
                    du[4 * k4] = fmaf(bf, w4.x, du[4 * k4]); du[4 * k4 + 1] = fmaf(bf, w4.y, du[4 * k4 + 1]);
                    du[4 * k4 + 2] = fmaf(bf, w4.z, du[4 * k4 + 2]); du[4 * k4 + 3] = fmaf(bf, w4.w, du[4 * k4 + 3]);
                }
            }
        }
        // ReLU + LayerNorm + Linear(H,H) backward
        float dn[TH], m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < TH; ++i) {
            const float ga = sm[SC_GA + i];
            const float pre = fmaf(ga, nrm[i], sm[SC_BE + i]);
            const float d_o = pre > 0.f ? gq * du[i] : 0.f;
            const float s_g = warp_sumf(d_o * nrm[i]), s_b = warp_sumf(d_o);
            if (lane == 0) { atomicAdd(sm + SC_ACC + TH + i, s_g); atomicAdd(sm + SC_ACC + 2 * TH + i, s_b); }
            dn[i] = d_o * ga;
            m1 += dn[i];
            m2 = fmaf(dn[i], nrm[i], m2);
        }
        m1 /= (float)TH;
        m2 /= (float)TH;
        float dx[TH];
#pragma unroll
        for (int k = 0; k < TH; ++k) dx[k] = 0.f;
#pragma unroll
        for (int i = 0; i < TH; ++i) {
            const float dy = rstd * (dn[i] - m1 - nrm[i] * m2);
            dn[i] = dy;
            const float sdy = warp_sumf(dy);
            if (lane == 0) atomicAdd(sm + SC_ACC + i, sdy);
            const float4 *wr = reinterpret_cast<const float4 *>(sm + SC_W0 + i * TH);
#pragma unroll
            for (int k4 = 0; k4 < TH / 4; ++k4) {
                const float4 w4 = wr[k4];
                dx[4 * k4] = fmaf(w4.x, dy, dx[4 * k4]); dx[4 * k4 + 1] = fmaf(w4.y, dy, dx[4 * k4 + 1]);
                dx[4 * k4 + 2] = fmaf(w4.z, dy, dx[4 * k4 + 2]); dx[4 * k4 + 3] = fmaf(w4.w, dy, dx[4 * k4 + 3]);
            }
        }
        {
            const float sg = warp_sumf(gq);
            if (lane == 0) atomicAdd(sm + SC_ACC + 3 * TH, sg);
        }
        if (live) {
            float4 *o1 = reinterpret_cast<float4 *>(dF + cell * TH), *o2 = reinterpret_cast<float4 *>(dY + cell * TH);
#pragma unroll
            for (int k4 = 0; k4 < TH / 4; ++k4) {
                if (!PRE) o1[k4] = make_float4(dx[4 * k4], dx[4 * k4 + 1], dx[4 * k4 + 2], dx[4 * k4 + 3]);
                o2[k4] = make_float4(dn[4 * k4], dn[4 * k4 + 1], dn[4 * k4 + 2], dn[4 * k4 + 3]);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 3 * TH + 1) {
        const float v = sm[SC_ACC + threadIdx.x];
        if (v != 0.f) {
            float *dst = threadIdx.x < TH ? gb0 + threadIdx.x : threadIdx.x < 2 * TH ? ggamma + (threadIdx.x - TH)
                       : threadIdx.x < 3 * TH ? gbeta + (threadIdx.x - 2 * TH) : gb2;
            atomicAdd(dst, v);
        }
    }
}

// ---- weight gradients: lanes own hidden units, accumulators live in registers for the whole kernel ----
// per-warp staging of one tile (floats): GU[TT][TK] | DY[TT][TH] | FF[TT][TH] | G[TT] | BITS[TT][4] | HEAD[TT]
#define WS_GU 0
#define WS_DY (WS_GU + TT * TK)
#define WS_FF (WS_DY + TT * TH)
#define WS_G (WS_FF + TT * TH)
#define WS_BITS (WS_G + TT)
#define WS_HEAD (WS_BITS + TT * 4)
#define WS_END (WS_HEAD + TT)
#define SW_V 0                            // block accumulator of V: [(k*4+jj)*32 + lane]
#define SW_P (SW_V + TK * TJ)             // P[TJ]
#define SW_W0 (SW_P + TJ)                 // dW0[TH][TH]
#define SW_END (SW_W0 + TH * TH)

__global__ void __launch_bounds__(TW * 32, 1)
k_tailc_bwd_weights(const int32_t *__restrict__ counters, int cap, const float *__restrict__ F, const float *__restrict__ O,
                    const float *__restrict__ dY, const float *__restrict__ Gc, const uint4 *__restrict__ bits,
                    const int32_t *__restrict__ cell_key, const int32_t *__restrict__ slot_head,
                    const float *__restrict__ rel, float *__restrict__ Vg, float *__restrict__ Pg, float *__restrict__ Phg,
                    float *__restrict__ gW0)
{
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long C = min(counters[0], cap);
    if ((long long)blockIdx.x * TW * TT >= C) return;
    for (int i = threadIdx.x; i < SW_END; i += TW * 32) sm[i] = 0.f;
    __syncthreads();
    float *ws = sm + SW_END + warp * WS_END;
    float4 V[TK];
#pragma unroll
    for (int k = 0; k < TK; ++k) V[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 P = make_float4(0.f, 0.f, 0.f, 0.f), Ph = P;
    float w0a[TH / 2];
#pragma unroll
    for (int k = 0; k < TH / 2; ++k) w0a[k] = 0.f;
    int cur_head = -1;
    const int i16 = lane & 15, half = lane >> 4;
    auto flush_head = [&]() {
        if (cur_head >= 0) {
            float *p = Phg + (size_t)cur_head * TJ + 4 * lane;
            if (Ph.x != 0.f) atomicAdd(p, Ph.x);
            if (Ph.y != 0.f) atomicAdd(p + 1, Ph.y);
            if (Ph.z != 0.f) atomicAdd(p + 2, Ph.z);
            if (Ph.w != 0.f) atomicAdd(p + 3, Ph.w);
        }
        Ph = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    for (long long tile = (long long)blockIdx.x * TW + warp; tile * TT < C; tile += (long long)gridDim.x * TW) {
        const long long cell0 = tile * TT;
        const int nc = (int)min((long long)TT, C - cell0);
        // ---- stage the tile: every load is contiguous over the tile's cells ----
        if (lane < TT) {
            const bool ok = lane < nc;
            ws[WS_G + lane] = ok ? Gc[cell0 + lane] : 0.f;
            reinterpret_cast<int *>(ws)[WS_HEAD + lane] = ok ? slot_head[cell_key[cell0 + lane] >> 5] : -1;
            reinterpret_cast<uint4 *>(ws + WS_BITS)[lane] = ok ? bits[cell0 + lane] : make_uint4(0u, 0u, 0u, 0u);
        }
        for (int v = lane; v < TT * TH / 4; v += 32) {           // 64 float4 per array
            const bool ok = (v >> 2) < nc;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            reinterpret_cast<float4 *>(ws + WS_DY)[v] = ok ? __ldg(reinterpret_cast<const float4 *>(dY + cell0 * TH) + v) : z4;
            reinterpret_cast<float4 *>(ws + WS_FF)[v] = ok ? __ldg(reinterpret_cast<const float4 *>(F + cell0 * TH) + v) : z4;
            const float4 o4 = ok ? __ldg(reinterpret_cast<const float4 *>(O + cell0 * TH) + v) : z4;
            reinterpret_cast<float4 *>(ws + WS_GU + (v >> 2) * TK)[v & 3] = o4;              // u[0..16) = o
        }
        __syncwarp();
        for (int v = lane; v < TT * TH; v += 32) {               // u[16..32) = rel[head], then scale the row by g
            const int cl = v >> 4, k = v & 15;
            const int head = reinterpret_cast<const int *>(ws)[WS_HEAD + cl];
            const float gq = ws[WS_G + cl];
            ws[WS_GU + cl * TK + TH + k] = head >= 0 ? gq * __ldg(rel + (size_t)head * TH + k) : 0.f;
            ws[WS_GU + cl * TK + k] *= gq;
        }
        __syncwarp();
        for (int c = 0; c < TT; ++c) {
            const int head = reinterpret_cast<const int *>(ws)[WS_HEAD + c];
            if (head != cur_head) {                              // warp-uniform: cells are ordered by slot
                flush_head();
                cur_head = head;
            }
            const uint32_t wbits = reinterpret_cast<const uint32_t *>(ws + WS_BITS)[c * 4 + (lane >> 3)] >> ((lane & 7) * 4);
            const float b0f = (float)(wbits & 1u), b1f = (float)((wbits >> 1) & 1u), b2f = (float)((wbits >> 2) & 1u),
                        b3f = (float)((wbits >> 3) & 1u);
            const float gq = ws[WS_G + c];
            P.x = fmaf(b0f, gq, P.x); P.y = fmaf(b1f, gq, P.y); P.z = fmaf(b2f, gq, P.z); P.w = fmaf(b3f, gq, P.w);
            Ph.x = fmaf(b0f, gq, Ph.x); Ph.y = fmaf(b1f, gq, Ph.y); Ph.z = fmaf(b2f, gq, Ph.z); Ph.w = fmaf(b3f, gq, Ph.w);
            const float4 *u4 = reinterpret_cast<const float4 *>(ws + WS_GU + c * TK);
#pragma unroll
            for (int k4 = 0; k4 < TK / 4; ++k4) {
                const float4 u = u4[k4];
                V[4 * k4].x = fmaf(b0f, u.x, V[4 * k4].x); V[4 * k4].y = fmaf(b1f, u.x, V[4 * k4].y);
                V[4 * k4].z = fmaf(b2f, u.x, V[4 * k4].z); V[4 * k4].w = fmaf(b3f, u.x, V[4 * k4].w);
                V[4 * k4 + 1].x = fmaf(b0f, u.y, V[4 * k4 + 1].x); V[4 * k4 + 1].y = fmaf(b1f, u.y, V[4 * k4 + 1].y);
                V[4 * k4 + 1].z = fmaf(b2f, u.y, V[4 * k4 + 1].z); V[4 * k4 + 1].w = fmaf(b3f, u.y, V[4 * k4 + 1].w);
                V[4 * k4 + 2].x = fmaf(b0f, u.z, V[4 * k4 + 2].x); V[4 * k4 + 2].y = fmaf(b1f, u.z, V[4 * k4 + 2].y);
                V[4 * k4 + 2].z = fmaf(b2f, u.z, V[4 * k4 + 2].z); V[4 * k4 + 2].w = fmaf(b3f, u.z, V[4 * k4 + 2].w);
                V[4 * k4 + 3].x = fmaf(b0f, u.w, V[4 * k4 + 3].x); V[4 * k4 + 3].y = fmaf(b1f, u.w, V[4 * k4 + 3].y);
                V[4 * k4 + 3].z = fmaf(b2f, u.w, V[4 * k4 + 3].z); V[4 * k4 + 3].w = fmaf(b3f, u.w, V[4 * k4 + 3].w);
            }
            // dW0[i][k] += dy_i * F_k on (unit i, half of k) lane pairs
            const float dyi = ws[WS_DY + c * TH + i16];
            const float4 *f4 = reinterpret_cast<const float4 *>(ws + WS_FF + c * TH + half * 8);
            const float4 fa = f4[0], fb = f4[1];
            w0a[0] = fmaf(dyi, fa.x, w0a[0]); w0a[1] = fmaf(dyi, fa.y, w0a[1]); w0a[2] = fmaf(dyi, fa.z, w0a[2]); w0a[3] = fmaf(dyi, fa.w, w0a[3]);
            w0a[4] = fmaf(dyi, fb.x, w0a[4]); w0a[5] = fmaf(dyi, fb.y, w0a[5]); w0a[6] = fmaf(dyi, fb.z, w0a[6]); w0a[7] = fmaf(dyi, fb.w, w0a[7]);
        }
        __syncwarp();
    }
    flush_head();
    // ---- once per kernel: warps add their registers into the block accumulators, then one atomic per element ----
    {
        float *dst = sm + SW_V + lane;
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            atomicAdd(dst + (k * 4 + 0) * 32, V[k].x);
            atomicAdd(dst + (k * 4 + 1) * 32, V[k].y);
            atomicAdd(dst + (k * 4 + 2) * 32, V[k].z);
            atomicAdd(dst + (k * 4 + 3) * 32, V[k].w);
        }
        atomicAdd(sm + SW_P + 4 * lane, P.x); atomicAdd(sm + SW_P + 4 * lane + 1, P.y);
        atomicAdd(sm + SW_P + 4 * lane + 2, P.z); atomicAdd(sm + SW_P + 4 * lane + 3, P.w);
#pragma unroll
        for (int k = 0; k < TH / 2; ++k) atomicAdd(sm + SW_W0 + i16 * TH + half * 8 + k, w0a[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TK * TJ; i += TW * 32) {       // sm index (k*4+jj)*32 + l  ->  V[4*l+jj][k]
        const int l = i & 31, jj = (i >> 5) & 3, k = i >> 7;
        const float v = sm[SW_V + i];
        if (v != 0.f) atomicAdd(Vg + (4 * l + jj) * TK + k, v);
    }
    for (int i = threadIdx.x; i < TJ; i += TW * 32) if (sm[SW_P + i] != 0.f) atomicAdd(Pg + i, sm[SW_P + i]);
    if (gW0)
        for (int i = threadIdx.x; i < TH * TH; i += TW * 32) if (sm[SW_W0 + i] != 0.f) atomicAdd(gW0 + i, sm[SW_W0 + i]);
}

// dW1, db1, dW2 from V and P (one block, thread j)
__global__ void __launch_bounds__(TJ)
k_tailc_finish(const float *__restrict__ W1, const float *__restrict__ b1, const float *__restrict__ W2,
               const float *__restrict__ Vg, const float *__restrict__ Pg, float *__restrict__ gW1,
               float *__restrict__ gb1, float *__restrict__ gW2)
{
    const int j = threadIdx.x;
    const float w2 = W2[j], p = Pg[j];
    float s = b1[j] * p;
    for (int k = 0; k < TK; ++k) {
        const float v = Vg[j * TK + k];
        gW1[j * TK + k] += w2 * v;
        s = fmaf(W1[j * TK + k], v, s);
    }
    gb1[j] += w2 * p;
    gW2[j] += s;
}

// relation-embedding gradient: du[16+k] summed over the cells of head q = sum_j W1[j][16+k] W2_j P_q[j]
__global__ void __launch_bounds__(TH)
k_tailc_rel_grad(const float *__restrict__ W1, const float *__restrict__ W2, const float *__restrict__ Phg,
                 float *__restrict__ grad_rel)
{
    const int head = blockIdx.x, k = threadIdx.x;
    const float *p = Phg + (size_t)head * TJ;
    float acc = 0.f;
    for (int j = 0; j < TJ; ++j) acc = fmaf(p[j] * W2[j], W1[j * TK + TH + k], acc);
    if (acc != 0.f) grad_rel[(size_t)head * TH + k] += acc;
}

static int sm_count()
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

/* floats of scratch rl_tail_backward needs (V, P, P per head relation) */
int64_t rl_tail_scratch_floats(int32_t R) { return (int64_t)TK * TJ + TJ + (int64_t)R * TJ; }

int rl_tail_forward(const rl_cells *c, const int32_t *slot_head, int32_t H, int32_t J, const float *F, const float *W0,
                    const float *b0, const float *gamma, const float *beta, const float *W1, const float *b1,
                    const float *W2, const float *b2, const float *rel_emb, float *zc, float *O, uint32_t *relu_bits,
                    int32_t front_done, void *stream)
{
    if (!c || !c->counters || !c->cell_key || !slot_head || !F || !W0 || !b0 || !gamma || !beta || !W1 || !b1 || !W2 || !b2 ||
        !rel_emb || !zc || !O || !relu_bits)
        return rl_fail(RL_ERR_ARG, "rl_tail_forward: null argument");
    if (H != TH || J != TJ) return rl_fail(RL_ERR_ARG, "rl_tail_forward: built for hidden_dim 16 and a 128-wide score MLP");
    TailP w{W0, b0, gamma, beta, W1, b1, W2, b2, rel_emb};
    const size_t smem = (size_t)SC_END * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_tailc_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_tailc_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    const long long want = ((long long)c->cap + 255) / 256;
    const int grid = (int)(want < 4LL * sm_count() ? want : 4LL * sm_count());
    if (front_done) k_tailc_fwd<true><<<grid, 256, smem, (cudaStream_t)stream>>>(c->counters, c->cap, F, c->cell_key, slot_head, w, zc, O,
                                                                                 reinterpret_cast<uint4 *>(relu_bits));
    else k_tailc_fwd<false><<<grid, 256, smem, (cudaStream_t)stream>>>(c->counters, c->cap, F, c->cell_key, slot_head, w, zc, O,
                                                                       reinterpret_cast<uint4 *>(relu_bits));
    CHECK_LAUNCH("k_tailc_fwd");
    return RL_OK;
}

/* grads are ACCUMULATED; O / relu_bits come from rl_tail_forward of the same cells; scratch: rl_tail_scratch_floats(R) */
int rl_tail_backward(const rl_cells *c, const int32_t *slot_head, int32_t R, int32_t H, int32_t J, const float *F,
                     const float *W0, const float *b0, const float *gamma, const float *beta, const float *W1,
                     const float *b1, const float *W2, const float *b2, const float *rel_emb, const float *Gc,
                     const float *O, const uint32_t *relu_bits, float *dF, float *dY, float *gW0, float *gb0,
                     float *ggamma, float *gbeta, float *gW1, float *gb1, float *gW2, float *gb2, float *grel,
                     float *scratch, int32_t front_done, void *stream)
{
    if (!c || !c->counters || !c->cell_key || !slot_head || !F || !W0 || !b0 || !gamma || !beta || !W1 || !b1 || !W2 || !b2 || !rel_emb ||
        !Gc || !O || !relu_bits || (!front_done && (!dF || !gW0)) || !dY || !gb0 || !ggamma || !gbeta || !gW1 || !gb1 || !gW2 || !gb2 || !grel || !scratch || R <= 0)
        return rl_fail(RL_ERR_ARG, "rl_tail_backward: null argument");
    if (H != TH || J != TJ) return rl_fail(RL_ERR_ARG, "rl_tail_backward: built for hidden_dim 16 and a 128-wide score MLP");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)rl_tail_scratch_floats(R) * sizeof(float), st);
    if (e != cudaSuccess) return rl_fail(RL_ERR_CUDA, "rl_tail_backward: memset", e);
    float *Vg = scratch, *Pg = Vg + TK * TJ, *Phg = Pg + TJ;
    TailP w{W0, b0, gamma, beta, W1, b1, W2, b2, rel_emb};
    const uint4 *bits4 = reinterpret_cast<const uint4 *>(relu_bits);
    const size_t smem_a = (size_t)SC_END * sizeof(float);
    const size_t smem_b = (size_t)(SW_END + TW * WS_END) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_tailc_bwd_cells<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
        cudaFuncSetAttribute(k_tailc_bwd_cells<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
        cudaFuncSetAttribute(k_tailc_bwd_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
        attr = true;
    }
    const long long want = ((long long)c->cap + 255) / 256;
    const int grid = (int)(want < 4LL * sm_count() ? want : 4LL * sm_count());
    if (front_done) k_tailc_bwd_cells<true><<<grid, 256, smem_a, st>>>(c->counters, c->cap, F, w, Gc, bits4, dY, dY, gb0, ggamma, gbeta, gb2);
    else k_tailc_bwd_cells<false><<<grid, 256, smem_a, st>>>(c->counters, c->cap, F, w, Gc, bits4, dF, dY, gb0, ggamma, gbeta, gb2);
    CHECK_LAUNCH("k_tailc_bwd_cells");
    k_tailc_bwd_weights<<<sm_count(), TW * 32, smem_b, st>>>(c->counters, c->cap, F, O, dY, Gc, bits4, c->cell_key, slot_head,
                                                              rel_emb, Vg, Pg, Phg, front_done ? nullptr : gW0);
    CHECK_LAUNCH("k_tailc_bwd_weights");
    k_tailc_finish<<<1, TJ, 0, st>>>(W1, b1, W2, Vg, Pg, gW1, gb1, gW2);
    CHECK_LAUNCH("k_tailc_finish");
    k_tailc_rel_grad<<<R, TH, 0, st>>>(W1, W2, Phg, grel);
    CHECK_LAUNCH("k_tailc_rel_grad");
    return RL_OK;
}

}  // extern "C"
