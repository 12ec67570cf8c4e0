// rnnlogic_b200 -- RotatE entity-feature epilogue (reference: src/embedding.py:28-70).
//
//   score[b][e] = gamma - sum_d | h_b o rot(r) - e |_d        (complex modulus per dimension)
//
// This is NOT a GEMM: per (query, entity, dimension) it costs two subtractions, two FMAs and one
// square root, so it is FP32-ALU / MUFU bound (13*B*N*D flop), and it never materialises the
// reference's [B*N, 2D] gathers (3.7 GB at FB15k-237, D = 1000).  Layouts are entity-major:
// P[S][2D][32] projected heads, out[S][N][32].
#include "rl_device.cuh"

#define RT_ENT 64   // entities per block (8 per warp)
#define RT_DT 32    // dimensions per shared-memory tile

// P[slot][d][lane] = re(h o rot), P[slot][D+d][lane] = im(h o rot); embedding.py:31-38,54-60
__global__ void __launch_bounds__(256)
k_rotate_project(int D, float inv_scale, const float *__restrict__ eemb, const float *__restrict__ remb,
                 const int32_t *__restrict__ slot_head, const int32_t *__restrict__ lane_h, float *__restrict__ P)
{
    const int slot = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // d * 32 + lane
    if (i >= D * RL_LANES) return;
    const int d = i >> 5, lane = i & 31;
    const int h = lane_h[slot * RL_LANES + lane];
    float pr = 0.f, pi = 0.f;
    if (h >= 0) {
        const float theta = remb[(size_t)slot_head[slot] * D + d] / inv_scale;     // vec / (range / pi)
        const float c = cosf(theta), s = sinf(theta);
        const float re = eemb[(size_t)h * 2 * D + d], im = eemb[(size_t)h * 2 * D + D + d];
        pr = re * c - im * s;
        pi = re * s + im * c;
    }
    P[((size_t)slot * 2 * D + d) * RL_LANES + lane] = pr;
    P[((size_t)slot * 2 * D + D + d) * RL_LANES + lane] = pi;
}

// |z| = z^2 * rsqrt(z^2): one MUFU + one FMUL instead of the IEEE sqrt sequence (error ~2 ulp per
// term, ~1e-7 relative on the sum over D terms -- far inside the 1e-5 parity bar).
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));     // one MUFU.RSQ, no denormal fix-up code
    return y;
}
__device__ __forceinline__ float fast_norm2(float dr, float di)
{
    const float n2 = fmaf(dr, dr, di * di);
    return n2 * rsqrt_approx(fmaxf(n2, 1e-30f));                // 0 at the origin
}

__global__ void __launch_bounds__(256)
k_rotate_scores(int N, int D, float gamma, const float *__restrict__ eemb, const float *__restrict__ P,
                float *__restrict__ out)
{
    __shared__ float2 sP[RT_DT][32];              // (re, im) of the projected heads, [d][query]
    __shared__ float2 sE[RT_ENT][RT_DT];          // (re, im) of the block's entities, [entity][d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.y;
    const int e0 = blockIdx.x * RT_ENT;
    const float *Ps = P + (size_t)slot * 2 * D * RL_LANES;
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0;
    for (int d0 = 0; d0 < D; d0 += RT_DT) {
        const int nd = min(RT_DT, D - d0);
        __syncthreads();
        for (int i = threadIdx.x; i < RT_DT * 32; i += 256) {               // projected heads, coalesced
            const int d = i >> 5, l = i & 31;
            float2 v = make_float2(0.f, 0.f);
            if (d < nd) { v.x = Ps[((size_t)d0 + d) * RL_LANES + l]; v.y = Ps[((size_t)D + d0 + d) * RL_LANES + l]; }
            sP[d][l] = v;
        }
        for (int i = threadIdx.x; i < RT_ENT * RT_DT; i += 256) {            // entity rows, 128-byte segments
            const int ent = i / RT_DT, d = i % RT_DT;
            const int e = e0 + ent;
            float2 v = make_float2(0.f, 0.f);
            if (e < N && d < nd) { v.x = eemb[(size_t)e * 2 * D + d0 + d]; v.y = eemb[(size_t)e * 2 * D + D + d0 + d]; }
            sE[ent][d] = v;
        }
        __syncthreads();
        float part_sum[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) part_sum[j] = 0.f;
#pragma unroll 4
        for (int d = 0; d < nd; ++d) {
            const float2 p = sP[d][lane];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 ev = sE[warp * 8 + j][d];                       // broadcast
                part_sum[j] += fast_norm2(p.x - ev.x, p.y - ev.y);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += (double)part_sum[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int e = e0 + warp * 8 + j;
        if (e < N) out[((size_t)slot * N + e) * RL_LANES + lane] = gamma - (float)acc[j];
    }
}

// Backward.  Block = (64 entities, 32 dimensions), lanes = dimension; loops over ALL slots so the
// entity-table gradient needs no atomics.  dP[S][2D][32] (gradient w.r.t. the projected heads) is
// reduced over the block's entities in shared memory, then one atomicAdd per (slot, d, lane, block).
__global__ void __launch_bounds__(256)
k_rotate_bwd(int N, int D, int S, const float *__restrict__ eemb, const float *__restrict__ P,
             const float *__restrict__ G, float *__restrict__ d_eemb, float *__restrict__ dP)
{
    __shared__ float sP[2][32][RT_DT + 1];        // [part][query][d]
    __shared__ float sG[RT_ENT][32];              // [entity][query]
    __shared__ float sR[2][32][RT_DT + 1];        // block partial of dP  [part][query][d] (shared atomics)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e0 = blockIdx.x * RT_ENT;
    const int d = blockIdx.y * RT_DT + lane;
    const bool dok = d < D;
    float er[8], ei[8], ger[8], gei[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int e = e0 + warp * 8 + j;
        er[j] = (e < N && dok) ? eemb[(size_t)e * 2 * D + d] : 0.f;
        ei[j] = (e < N && dok) ? eemb[(size_t)e * 2 * D + D + d] : 0.f;
        ger[j] = 0.f;
        gei[j] = 0.f;
    }
    for (int slot = 0; slot < S; ++slot) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * RT_DT * 32; i += 256) {
            const int part = i / (RT_DT * 32), rem = i % (RT_DT * 32), dd = rem >> 5, l = rem & 31;
            const int dg = blockIdx.y * RT_DT + dd;
            sP[part][l][dd] = dg < D ? P[((size_t)slot * 2 * D + (size_t)part * D + dg) * RL_LANES + l] : 0.f;
        }
        for (int i = threadIdx.x; i < RT_ENT * 32; i += 256) {
            const int ent = i >> 5, l = i & 31;
            const int e = e0 + ent;
            sG[ent][l] = e < N ? G[((size_t)slot * N + e) * RL_LANES + l] : 0.f;
        }
        for (int i = threadIdx.x; i < 2 * 32 * (RT_DT + 1); i += 256) (&sR[0][0][0])[i] = 0.f;
        __syncthreads();
        for (int b = 0; b < 32; ++b) {
            const float pr = sP[0][b][lane], pi = sP[1][b][lane];
            float dpr = 0.f, dpi = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float g = sG[warp * 8 + j][b];
                const float dr = pr - er[j], di = pi - ei[j];
                const float n2 = dr * dr + di * di;
                const float t = n2 > 0.f ? g * rsqrt_approx(fmaxf(n2, 1e-30f)) : 0.f;   // d|z|/dz = z/|z|, 0 at the origin (torch.norm)
                ger[j] += t * dr;                                     // score = gamma - sum |p - e|
                gei[j] += t * di;
                dpr -= t * dr;
                dpi -= t * di;
            }
            atomicAdd(&sR[0][b][lane], dpr);
            atomicAdd(&sR[1][b][lane], dpi);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * 32 * RT_DT; i += 256) {
            const int part = i / (32 * RT_DT), rem = i % (32 * RT_DT), dd = rem >> 5, l = rem & 31;   // l = query
            const int dg = blockIdx.y * RT_DT + dd;
            if (dg < D) {
                const float v = sR[part][l][dd];
                if (v != 0.f) atomicAdd(dP + ((size_t)slot * 2 * D + (size_t)part * D + dg) * RL_LANES + l, v);
            }
        }
    }
    if (dok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = e0 + warp * 8 + j;
            if (e < N) {
                atomicAdd(d_eemb + (size_t)e * 2 * D + d, ger[j]);         // d_eemb also receives the head part below
                atomicAdd(d_eemb + (size_t)e * 2 * D + D + d, gei[j]);
            }
        }
    }
}

// chain rule through p = h o (cos theta + i sin theta), theta = remb[q] / (range/pi)
__global__ void __launch_bounds__(256)
k_rotate_project_bwd(int D, float inv_scale, const float *__restrict__ eemb, const float *__restrict__ remb,
                     const int32_t *__restrict__ slot_head, const int32_t *__restrict__ lane_h,
                     const float *__restrict__ dP, float *__restrict__ d_eemb, float *__restrict__ d_remb)
{
    const int slot = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = i >> 5, lane = i & 31;
    float dtheta = 0.f;
    const int q = slot_head[slot];
    if (d < D) {
        const int h = lane_h[slot * RL_LANES + lane];
        if (h >= 0) {
            const float theta = remb[(size_t)q * D + d] / inv_scale;
            const float c = cosf(theta), s = sinf(theta);
            const float re = eemb[(size_t)h * 2 * D + d], im = eemb[(size_t)h * 2 * D + D + d];
            const float gr = dP[((size_t)slot * 2 * D + d) * RL_LANES + lane];
            const float gi = dP[((size_t)slot * 2 * D + D + d) * RL_LANES + lane];
            atomicAdd(d_eemb + (size_t)h * 2 * D + d, gr * c + gi * s);
            atomicAdd(d_eemb + (size_t)h * 2 * D + D + d, -gr * s + gi * c);
            dtheta = gr * (-re * s - im * c) + gi * (re * c - im * s);
        }
    }
    dtheta = warp_sumf(dtheta);                                    // the 32 lanes of a warp share (slot, d)
    if (lane == 0 && d < D && dtheta != 0.f) atomicAdd(d_remb + (size_t)q * D + d, dtheta / inv_scale);
}

extern "C" {

int rl_rotate_scores(const rl_graph *g, const rl_slots *s, int32_t D, float gamma, const float *eemb,
                     const float *remb, float *P, float *out, void *stream)
{
    if (!g || !s || !eemb || !remb || !P || !out || D <= 0) return rl_fail(RL_ERR_ARG, "rl_rotate_scores: bad argument");
    const int S = s->num_slots, N = g->num_entities;
    if (S <= 0) return RL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const float inv_scale = (float)(((double)gamma + 2.0) / D / 3.141592653589793238462643383279);
    k_rotate_project<<<dim3((D * RL_LANES + 255) / 256, S), 256, 0, st>>>(D, inv_scale, eemb, remb, s->slot_head, s->lane_h, P);
    CHECK_LAUNCH("k_rotate_project");
    k_rotate_scores<<<dim3((N + RT_ENT - 1) / RT_ENT, S), 256, 0, st>>>(N, D, gamma, eemb, P, out);
    CHECK_LAUNCH("k_rotate_scores");
    return RL_OK;
}

int rl_rotate_backward(const rl_graph *g, const rl_slots *s, int32_t D, float gamma, const float *eemb,
                       const float *remb, const float *P, const float *G, float *dP, float *d_eemb,
                       float *d_remb, void *stream)
{
    if (!g || !s || !eemb || !remb || !P || !G || !dP || !d_eemb || !d_remb || D <= 0)
        return rl_fail(RL_ERR_ARG, "rl_rotate_backward: bad argument");
    const int S = s->num_slots, N = g->num_entities;
    if (S <= 0) return RL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const float inv_scale = (float)(((double)gamma + 2.0) / D / 3.141592653589793238462643383279);
    k_rotate_bwd<<<dim3((N + RT_ENT - 1) / RT_ENT, (D + RT_DT - 1) / RT_DT), 256, 0, st>>>(N, D, S, eemb, P, G, d_eemb, dP);
    CHECK_LAUNCH("k_rotate_bwd");
    k_rotate_project_bwd<<<dim3((D * RL_LANES + 255) / 256, S), 256, 0, st>>>(D, inv_scale, eemb, remb, s->slot_head, s->lane_h, dP, d_eemb, d_remb);
    CHECK_LAUNCH("k_rotate_project_bwd");
    return RL_OK;
}

}  // extern "C"
