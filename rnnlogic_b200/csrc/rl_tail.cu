// rnnlogic_b200 -- fused dense tail of PredictorPlus with the `sum` aggregator
// (reference: src/layers.py:73-75 + src/predictors.py:253-255):
//
//   y = W0 F + b0 ; o = relu(LayerNorm(y)) ; u = [o, relation_emb[q]] ;
//   z = W2 . relu(W1 u + b1) + b2                         (one scalar per candidate cell)
//
// One thread per candidate, weights broadcast from shared memory, everything in fp32 FFMA (the
// 1e-5 parity bar rules out TF32; 4.5 kFMA per candidate is nowhere near a bound).  No [C,128]
// activation is materialised in the forward.  The backward recomputes the forward, reduces the
// small gradients (b0, gamma, beta, b1, W2, b2) in-kernel and writes the per-candidate factors
// (delta1[C][J], u[C][2H], dy[C][H], d_rel[C][H]) of the three outer-product gradients
// (W1, W0, relation_emb), which the caller contracts with one GEMM / index_add each.
#include "rl_device.cuh"

struct TailW {
    const float *W0, *b0, *gamma, *beta, *W1, *b1, *W2, *b2, *rel;   // rel = relation_emb.weight [R][H]
    int J;                                                            // hidden width of the score MLP (128)
};

template <int H>
__device__ __forceinline__ void tail_front(const float *__restrict__ sW0, const float *__restrict__ sb0,
                                           const float *__restrict__ sg, const float *__restrict__ sbt,
                                           const float *x, float *y, float *nrm, float *u, float &rstd)
{
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        float a = sb0[i];
#pragma unroll
        for (int k = 0; k < H; ++k) a = fmaf(sW0[i * H + k], x[k], a);
        y[i] = a;
        mean += a;
    }
    mean /= (float)H;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < H; ++i) { const float d = y[i] - mean; var = fmaf(d, d, var); }
    var /= (float)H;
    rstd = rsqrtf(var + 1e-5f);
#pragma unroll
    for (int i = 0; i < H; ++i) {
        nrm[i] = (y[i] - mean) * rstd;
        u[i] = fmaxf(fmaf(sg[i], nrm[i], sbt[i]), 0.f);
    }
}

template <int H>
__global__ void __launch_bounds__(256)
k_sum_tail_fwd(long long C, const float *__restrict__ F, const int32_t *__restrict__ cand_head, TailW w,
               float *__restrict__ z)
{
    extern __shared__ float sm[];
    const int J = w.J;
    float *sW0 = sm, *sb0 = sW0 + H * H, *sg = sb0 + H, *sbt = sg + H, *sW1 = sbt + H, *sb1 = sW1 + J * 2 * H, *sW2 = sb1 + J;
    for (int i = threadIdx.x; i < H * H; i += 256) sW0[i] = w.W0[i];
    for (int i = threadIdx.x; i < H; i += 256) { sb0[i] = w.b0[i]; sg[i] = w.gamma[i]; sbt[i] = w.beta[i]; }
    for (int i = threadIdx.x; i < J * 2 * H; i += 256) sW1[i] = w.W1[i];
    for (int i = threadIdx.x; i < J; i += 256) { sb1[i] = w.b1[i]; sW2[i] = w.W2[i]; }
    __syncthreads();
    const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    float x[H], y[H], nrm[H], u[2 * H], rstd;
#pragma unroll
    for (int k = 0; k < H; ++k) x[k] = F[c * H + k];
    tail_front<H>(sW0, sb0, sg, sbt, x, y, nrm, u, rstd);
    const float *rv = w.rel + (size_t)cand_head[c] * H;
#pragma unroll
    for (int k = 0; k < H; ++k) u[H + k] = __ldg(rv + k);
    float acc = w.b2[0];
    for (int j = 0; j < J; ++j) {
        float a = sb1[j];
        const float *wr = sW1 + j * 2 * H;
#pragma unroll
        for (int k = 0; k < 2 * H; ++k) a = fmaf(wr[k], u[k], a);
        acc = fmaf(sW2[j], fmaxf(a, 0.f), acc);
    }
    z[c] = acc;
}

template <int H>
__global__ void __launch_bounds__(256)
k_sum_tail_bwd(long long C, const float *__restrict__ F, const int32_t *__restrict__ cand_head, TailW w,
               const float *__restrict__ dz, float *__restrict__ dF, float *__restrict__ delta1,
               float *__restrict__ U, float *__restrict__ dY, float *__restrict__ dRel,
               float *__restrict__ g_small)   // [b0 H | gamma H | beta H | b1 J | W2 J | b2 1]
{
    extern __shared__ float sm[];
    const int J = w.J;
    float *sW0 = sm, *sb0 = sW0 + H * H, *sg = sb0 + H, *sbt = sg + H, *sW1 = sbt + H, *sb1 = sW1 + J * 2 * H, *sW2 = sb1 + J;
    float *sacc = sW2 + J;                                            // block accumulators, same layout as g_small
    const int n_small = 3 * H + 2 * J + 1;
    for (int i = threadIdx.x; i < H * H; i += 256) sW0[i] = w.W0[i];
    for (int i = threadIdx.x; i < H; i += 256) { sb0[i] = w.b0[i]; sg[i] = w.gamma[i]; sbt[i] = w.beta[i]; }
    for (int i = threadIdx.x; i < J * 2 * H; i += 256) sW1[i] = w.W1[i];
    for (int i = threadIdx.x; i < J; i += 256) { sb1[i] = w.b1[i]; sW2[i] = w.W2[i]; }
    for (int i = threadIdx.x; i < n_small; i += 256) sacc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
    const bool live = c < C;
    float x[H], y[H], nrm[H], u[2 * H], du[2 * H], rstd = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) x[k] = live ? F[c * H + k] : 0.f;
    tail_front<H>(sW0, sb0, sg, sbt, x, y, nrm, u, rstd);
    const float *rv = w.rel + (size_t)(live ? cand_head[c] : 0) * H;
#pragma unroll
    for (int k = 0; k < H; ++k) u[H + k] = __ldg(rv + k);
#pragma unroll
    for (int k = 0; k < 2 * H; ++k) du[k] = 0.f;
    const float g = live ? dz[c] : 0.f;
    for (int j = 0; j < J; ++j) {
        float a = sb1[j];
        const float *wr = sW1 + j * 2 * H;
#pragma unroll
        for (int k = 0; k < 2 * H; ++k) a = fmaf(wr[k], u[k], a);
        const float act = fmaxf(a, 0.f);
        const float d1 = a > 0.f ? g * sW2[j] : 0.f;
        if (live) delta1[c * J + j] = d1;
#pragma unroll
        for (int k = 0; k < 2 * H; ++k) du[k] = fmaf(d1, wr[k], du[k]);
        const float s_w2 = warp_sumf(g * act), s_b1 = warp_sumf(d1);
        if (lane == 0) {
            atomicAdd(sacc + 3 * H + J + j, s_w2);
            atomicAdd(sacc + 3 * H + j, s_b1);
        }
    }
    {
        const float s = warp_sumf(g);
        if (lane == 0) atomicAdd(sacc + 3 * H + 2 * J, s);
    }
    // LayerNorm + ReLU + Linear(H,H) backward
    float dn[H], m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const float pre = fmaf(sg[i], nrm[i], sbt[i]);
        const float d_o = pre > 0.f ? du[i] : 0.f;
        const float s_g = warp_sumf(d_o * nrm[i]), s_b = warp_sumf(d_o);
        if (lane == 0) { atomicAdd(sacc + H + i, s_g); atomicAdd(sacc + 2 * H + i, s_b); }
        dn[i] = d_o * sg[i];
        m1 += dn[i];
        m2 = fmaf(dn[i], nrm[i], m2);
    }
    m1 /= (float)H;
    m2 /= (float)H;
    float dy[H], dx[H];
#pragma unroll
    for (int k = 0; k < H; ++k) dx[k] = 0.f;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        dy[i] = rstd * (dn[i] - m1 - nrm[i] * m2);
        const float s = warp_sumf(dy[i]);
        if (lane == 0) atomicAdd(sacc + i, s);
#pragma unroll
        for (int k = 0; k < H; ++k) dx[k] = fmaf(sW0[i * H + k], dy[i], dx[k]);
    }
    if (live) {
#pragma unroll
        for (int k = 0; k < H; ++k) {
            dF[c * H + k] = dx[k];
            dY[c * H + k] = dy[k];
            dRel[c * H + k] = du[H + k];
        }
#pragma unroll
        for (int k = 0; k < 2 * H; ++k) U[c * 2 * H + k] = u[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_small; i += 256)
        if (sacc[i] != 0.f) atomicAdd(g_small + i, sacc[i]);
}

extern "C" {

static size_t tail_smem(int H, int J, bool bwd)
{
    size_t n = (size_t)H * H + 3 * H + (size_t)J * 2 * H + 2 * J;
    if (bwd) n += 3 * H + 2 * J + 1;
    return n * sizeof(float);
}

int rl_sum_tail_forward(int64_t C, int32_t H, int32_t J, const float *F, const int32_t *cand_head, const float *W0,
                        const float *b0, const float *gamma, const float *beta, const float *W1, const float *b1,
                        const float *W2, const float *b2, const float *rel_emb, float *z, void *stream)
{
    if (!F || !cand_head || !W0 || !b0 || !gamma || !beta || !W1 || !b1 || !W2 || !b2 || !rel_emb || !z)
        return rl_fail(RL_ERR_ARG, "rl_sum_tail_forward: null argument");
    if ((H != 16 && H != 32) || J <= 0 || J > 256) return rl_fail(RL_ERR_ARG, "rl_sum_tail_forward: unsupported H (16|32) or J (<=256)");
    if (C <= 0) return RL_OK;
    TailW w{W0, b0, gamma, beta, W1, b1, W2, b2, rel_emb, J};
    const unsigned grid = (unsigned)((C + 255) / 256);
    const size_t smem = tail_smem(H, J, false);
    cudaStream_t st = (cudaStream_t)stream;
    if (H == 16) {
        cudaFuncSetAttribute(k_sum_tail_fwd<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_sum_tail_fwd<16><<<grid, 256, smem, st>>>(C, F, cand_head, w, z);
    } else {
        cudaFuncSetAttribute(k_sum_tail_fwd<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_sum_tail_fwd<32><<<grid, 256, smem, st>>>(C, F, cand_head, w, z);
    }
    CHECK_LAUNCH("k_sum_tail_fwd");
    return RL_OK;
}

int rl_sum_tail_backward(int64_t C, int32_t H, int32_t J, const float *F, const int32_t *cand_head, const float *W0,
                         const float *b0, const float *gamma, const float *beta, const float *W1, const float *b1,
                         const float *W2, const float *b2, const float *rel_emb, const float *dz, float *dF,
                         float *delta1, float *U, float *dY, float *dRel, float *g_small, void *stream)
{
    if (!F || !cand_head || !W0 || !b0 || !gamma || !beta || !W1 || !b1 || !W2 || !b2 || !rel_emb || !dz || !dF ||
        !delta1 || !U || !dY || !dRel || !g_small)
        return rl_fail(RL_ERR_ARG, "rl_sum_tail_backward: null argument");
    if ((H != 16 && H != 32) || J <= 0 || J > 256) return rl_fail(RL_ERR_ARG, "rl_sum_tail_backward: unsupported H (16|32) or J (<=256)");
    if (C <= 0) return RL_OK;
    TailW w{W0, b0, gamma, beta, W1, b1, W2, b2, rel_emb, J};
    const unsigned grid = (unsigned)((C + 255) / 256);
    const size_t smem = tail_smem(H, J, true);
    cudaStream_t st = (cudaStream_t)stream;
    if (H == 16) {
        cudaFuncSetAttribute(k_sum_tail_bwd<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_sum_tail_bwd<16><<<grid, 256, smem, st>>>(C, F, cand_head, w, dz, dF, delta1, U, dY, dRel, g_small);
    } else {
        cudaFuncSetAttribute(k_sum_tail_bwd<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_sum_tail_bwd<32><<<grid, 256, smem, st>>>(C, F, cand_head, w, dz, dF, delta1, U, dY, dRel, g_small);
    }
    CHECK_LAUNCH("k_sum_tail_bwd");
    return RL_OK;
}

}  // extern "C"
