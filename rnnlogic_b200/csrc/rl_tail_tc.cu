// rnnlogic_b200 -- dense tail of PredictorPlus on the candidate cells, with the score MLP on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in TMEM).
// Reference: src/layers.py:73-75 (Linear(H,H) -> LayerNorm -> ReLU) + src/predictors.py:253-255
// ([.., relation_emb[q]] -> Linear(2H,128) -> ReLU -> Linear(128,1)), H = 16, J = 128.
//
//   y = W0 F + b0 ; o = relu(LN(y)) ; u = [o, rel[q]] ; a = W1 u + b1 ; z = W2 . relu(a) + b2
//
// One THREAD per cell does the small front (16x16 Linear, LayerNorm) in registers and writes its row of the MMA
// operand straight into shared memory in the canonical UMMA layout (rl_umma.cuh) -- the operands are functions of the
// cell, so there is nothing for TMA to copy.  A tile is 128 cells = the M of one tcgen05.mma.
//
// The parity bar is 1e-5 on the logits, so plain TF32/BF16 inputs are out.  Both GEMM families are made exact enough
// by splitting:
//   forward  a[128 cells x 128] = U[128 x 32] W1^T : "3xTF32" -- U = Uhi + Ulo, W1 = Whi + Wlo (tf32 pieces),
//            Ulo Whi + Uhi Wlo + Uhi Whi accumulated in fp32 in TMEM (12 MMAs of K = 8), relative error ~2^-21.
//   backward the hidden layer is never recomputed: the forward leaves the 128 ReLU bits of a cell, and every
//            product of the backward has the 0/1 BIT matrix as one operand -- exact in bf16 -- so the other
//            operand is split into three bf16 pieces (3 x 8 significand bits = fp32) laid side by side along N:
//              du [128 cells x 32]   = BITS [cells x 128 j] (W2_j W1[j][k])              N = 3 x 32
//              V  [128 j x 33]      += BITS^T [j x cells]   (g_c [u_c, 1])                N = 3 x 40 (padded to 128)
//            The SAME shared-memory bit tile is the K-major A operand of the first and the MN-major A operand of the
//            second product.  V stays in TMEM for the whole kernel; everything else follows from it:
//              dW1[j][k] = W2_j V[j][k]   db1[j] = W2_j V[j][32]   dW2[j] = b1_j V[j][32] + sum_k W1[j][k] V[j][k]
// The LayerNorm / Linear(H,H) backward, dF, and the small bias gradients stay on the CUDA cores (per-thread
// accumulators, one reduction per kernel).
#include "rl_device.cuh"
#include "rl_umma.cuh"

#define TH 16         // hidden_dim
#define TJ 128        // hidden width of the score MLP
#define TK 32         // its input width (2H)
#define TILE 128      // cells per tile = MMA M
#define CHUNK_B (TILE * 16)   // bytes of one 16-byte chunk column of a 128-row tile

struct TailP { const float *W0, *b0, *gamma, *beta, *W1, *b1, *W2, *b2, *rel; };

// small parameters in shared memory (floats)
#define SP_W0 0
#define SP_B0 (SP_W0 + TH * TH)
#define SP_GA (SP_B0 + TH)
#define SP_BE (SP_GA + TH)
#define SP_B1 (SP_BE + TH)
#define SP_W2 (SP_B1 + TJ)
#define SP_END (SP_W2 + TJ)

__device__ __forceinline__ void load_small(float *sp, const TailP &w, bool pre)
{
    for (int i = threadIdx.x; i < TH * TH; i += blockDim.x) sp[SP_W0 + i] = pre ? 0.f : w.W0[i];
    for (int i = threadIdx.x; i < TH; i += blockDim.x) { sp[SP_B0 + i] = pre ? 0.f : w.b0[i]; sp[SP_GA + i] = w.gamma[i]; sp[SP_BE + i] = w.beta[i]; }
    for (int i = threadIdx.x; i < TJ; i += blockDim.x) { sp[SP_B1 + i] = w.b1[i]; sp[SP_W2 + i] = w.W2[i]; }
}

// front of one cell in registers: y = W0 f + b0 (PRE: f already is y), LayerNorm; returns 1/std, fills nrm[]
template <bool PRE>
__device__ __forceinline__ float cell_front(const float *sp, const float (&f)[TH], float (&nrm)[TH])
{
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < TH; ++i) {
        float a;
        if (PRE) a = f[i];
        else {
            a = sp[SP_B0 + i];
            const float4 *wr = reinterpret_cast<const float4 *>(sp + SP_W0 + i * TH);
#pragma unroll
            for (int k4 = 0; k4 < TH / 4; ++k4) {
                const float4 w4 = wr[k4];
                a = fmaf(w4.x, f[4 * k4], a); a = fmaf(w4.y, f[4 * k4 + 1], a);
                a = fmaf(w4.z, f[4 * k4 + 2], a); a = fmaf(w4.w, f[4 * k4 + 3], a);
            }
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

// ================================================================================================
// forward
// ================================================================================================
// shared memory (bytes): U hi | U lo | W1 hi | W1 lo  (8 chunks of 4 tf32 x 128 rows each) | small parameters | barrier
#define FS_UHI 0
#define FS_ULO (FS_UHI + 8 * CHUNK_B)
#define FS_WHI (FS_ULO + 8 * CHUNK_B)
#define FS_WLO (FS_WHI + 8 * CHUNK_B)
#define FS_SP (FS_WLO + 8 * CHUNK_B)
#define FS_BAR (FS_SP + SP_END * 4)
#define FS_END (FS_BAR + 16)

template <bool PRE>
__global__ void __launch_bounds__(TILE, 3)
k_tail_fwd_tc(const int32_t *__restrict__ counters, int cap, const float *__restrict__ F, const int32_t *__restrict__ cell_key,
              const int32_t *__restrict__ slot_head, TailP w, float *__restrict__ zc, uint4 *__restrict__ bits)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const long long C = min(counters[0], cap);
    const long long tiles = (C + TILE - 1) / TILE;
    if ((long long)blockIdx.x >= tiles) return;
    const int tid = threadIdx.x, warp = tid >> 5;
    float *sp = reinterpret_cast<float *>(smem + FS_SP);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + FS_BAR);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(smem + FS_BAR + 8);
    load_small(sp, w, PRE);
    {   // W1 row j = tid, split into tf32 pieces: B operand [n = j][k], K-major
        const float4 *wr = reinterpret_cast<const float4 *>(w.W1 + (size_t)tid * TK);
#pragma unroll
        for (int c = 0; c < TK / 4; ++c) {
            const float4 v = __ldg(wr + c);
            float4 hi, lo;
            umma::split_tf32(v.x, hi.x, lo.x); umma::split_tf32(v.y, hi.y, lo.y);
            umma::split_tf32(v.z, hi.z, lo.z); umma::split_tf32(v.w, hi.w, lo.w);
            *reinterpret_cast<float4 *>(smem + FS_WHI + c * CHUNK_B + tid * 16) = hi;
            *reinterpret_cast<float4 *>(smem + FS_WLO + c * CHUNK_B + tid * 16) = lo;
        }
    }
    if (tid == 0) umma::mbar_init(bar, 1);
    if (warp == 0) umma::tmem_alloc(tslot, 128);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tacc = *tslot;
    const uint32_t trow = tacc + ((uint32_t)(warp * 32) << 16);
    const float b2 = __ldg(w.b2);
    const uint32_t s_uhi = umma::smem_u32(smem + FS_UHI), s_ulo = umma::smem_u32(smem + FS_ULO);
    const uint32_t s_whi = umma::smem_u32(smem + FS_WHI), s_wlo = umma::smem_u32(smem + FS_WLO);
    constexpr uint32_t ID = umma::idesc(UMMA_FMT_TF32, TILE, TJ, false, false);
    uint32_t phase = 0;
    // the rows of a tile (F row, relation row of the cell's head) are fetched ONE TILE AHEAD: the dependent chain
    // cell_key -> slot_head -> rel row and the F row fly while the MMAs of the current tile run and its epilogue is computed
    float f[TH], r16[TH];
    auto fetch = [&](long long tile) -> bool {
        const long long cell = tile * TILE + tid;
        const bool ok = tile < tiles && cell < C;
        if (ok) {
            load_row16(F + cell * TH, f);
            load_row16(w.rel + (size_t)slot_head[cell_key[cell] >> 5] * TH, r16);
        } else {
#pragma unroll
            for (int k = 0; k < TH; ++k) { f[k] = 0.f; r16[k] = 0.f; }
        }
        return ok;
    };
    bool live = fetch(blockIdx.x);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long cell = tile * TILE + tid;
        float u[TK];
        {
            float nrm[TH];
            cell_front<PRE>(sp, f, nrm);
#pragma unroll
            for (int i = 0; i < TH; ++i) u[i] = live ? fmaxf(fmaf(sp[SP_GA + i], nrm[i], sp[SP_BE + i]), 0.f) : 0.f;
#pragma unroll
            for (int i = 0; i < TH; ++i) u[TH + i] = r16[i];
        }
#pragma unroll
        for (int c = 0; c < TK / 4; ++c) {
            float4 hi, lo;
            umma::split_tf32(u[4 * c], hi.x, lo.x); umma::split_tf32(u[4 * c + 1], hi.y, lo.y);
            umma::split_tf32(u[4 * c + 2], hi.z, lo.z); umma::split_tf32(u[4 * c + 3], hi.w, lo.w);
            *reinterpret_cast<float4 *>(smem + FS_UHI + c * CHUNK_B + tid * 16) = hi;
            *reinterpret_cast<float4 *>(smem + FS_ULO + c * CHUNK_B + tid * 16) = lo;
        }
        umma::fence_smem_to_async();
        umma::fence_before();                                  // this thread's TMEM reads of the previous tile are done
        __syncthreads();
        if (tid == 0) {
            umma::fence_after();
            // small products first: Ulo Whi, Uhi Wlo, then Uhi Whi; one K step = 8 tf32 = two chunks
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const uint32_t sa = p == 0 ? s_ulo : s_uhi, sb = p == 1 ? s_wlo : s_whi;
#pragma unroll
                for (int k = 0; k < TK / 8; ++k)
                    umma::mma_tf32(tacc, umma::desc(sa + k * 2 * CHUNK_B, CHUNK_B, 128), umma::desc(sb + k * 2 * CHUNK_B, CHUNK_B, 128),
                                   ID, (p | k) != 0);
            }
            umma::commit(bar);
        }
        const bool live_next = fetch(tile + gridDim.x);          // issued before the wait: in flight during the MMAs and the epilogue
        umma::mbar_wait(bar, phase);
        phase ^= 1u;
        umma::fence_after();
        float z0 = b2, z1 = 0.f;                                 // two chains: the 128-term dot product is latency-bound otherwise
        uint32_t word[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float a[32];
            umma::tmem_ld32(trow + q * 32, a);
            uint32_t wb = 0u;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float v0 = a[i] + sp[SP_B1 + q * 32 + i], v1 = a[i + 1] + sp[SP_B1 + q * 32 + i + 1];
                z0 = fmaf(sp[SP_W2 + q * 32 + i], fmaxf(v0, 0.f), z0);
                z1 = fmaf(sp[SP_W2 + q * 32 + i + 1], fmaxf(v1, 0.f), z1);
                wb |= (v0 > 0.f ? 1u : 0u) << i;
                wb |= (v1 > 0.f ? 1u : 0u) << (i + 1);
            }
            word[q] = wb;
        }
        if (live) {
            zc[cell] = z0 + z1;
            bits[cell] = make_uint4(word[0], word[1], word[2], word[3]);
        }
        live = live_next;
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_free(tacc, 128);
}

// ================================================================================================
// backward
// ================================================================================================
// shared memory (bytes)
#define BS_BITS 0                              // [16 chunks of 8 j][128 cells][16 B]    bf16 0/1
#define BS_GU (BS_BITS + 16 * CHUNK_B)         // [16 chunks of 8 n][128 cells][16 B]    bf16 pieces of g*[u,1]: n = 40*piece + k
#define BS_W21 (BS_GU + 16 * CHUNK_B)          // [16 chunks of 8 j][96 n][16 B]         bf16 pieces of W2_j W1[j][k]: n = 32*piece + k
#define BS_LUT (BS_W21 + 16 * 96 * 16)         // [256][16 B]  byte -> eight bf16 0/1
#define BS_SP (BS_LUT + 256 * 16)              // small parameters
#define BS_STG (BS_SP + SP_END * 4)            // per warp: dy[16][20] | f[16][20] floats (dW0 outer products, half a warp at a time)
#define STG_STRIDE 20
#define BS_ACC (BS_STG + 4 * 2 * 16 * STG_STRIDE * 4)   // dW0[256] | db0[16] dgamma[16] dbeta[16] db2[1]
#define BS_BAR (BS_ACC + (TH * TH + 3 * TH + 4) * 4)
#define BS_END (BS_BAR + 16)
#define GU_PIECE 40                            // columns per bf16 piece of g*[u,1] (33 used)

template <bool PRE>
__global__ void __launch_bounds__(TILE, 2)
k_tail_bwd_tc(const int32_t *__restrict__ counters, int cap, const float *__restrict__ F, TailP w,
              const float *__restrict__ Gc, const uint4 *__restrict__ bits, const int32_t *__restrict__ cell_key,
              const int32_t *__restrict__ slot_head, float *__restrict__ dF, float *__restrict__ dY,
              float *__restrict__ Vg, float *__restrict__ gW0, float *__restrict__ gb0, float *__restrict__ ggamma,
              float *__restrict__ gbeta, float *__restrict__ gb2, float *__restrict__ grel)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const long long C = min(counters[0], cap);
    const long long tiles = (C + TILE - 1) / TILE;
    // contiguous tile range per block: the cells are ordered by slot, so a block sees few head relations
    const long long t0 = tiles * blockIdx.x / gridDim.x, t1 = tiles * (blockIdx.x + 1) / gridDim.x;
    if (t0 >= t1) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *sp = reinterpret_cast<float *>(smem + BS_SP);
    float *acc = reinterpret_cast<float *>(smem + BS_ACC);
    float *stg = reinterpret_cast<float *>(smem + BS_STG) + warp * 2 * 16 * STG_STRIDE;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + BS_BAR);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(smem + BS_BAR + 8);
    load_small(sp, w, PRE);
    for (int i = tid; i < TH * TH + 3 * TH + 4; i += TILE) acc[i] = 0.f;
    for (int b = tid; b < 256; b += TILE) {
        uint32_t q[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) q[e] = ((b >> (2 * e)) & 1 ? 0x3F80u : 0u) | ((b >> (2 * e + 1)) & 1 ? 0x3F800000u : 0u);
        *reinterpret_cast<uint4 *>(smem + BS_LUT + b * 16) = make_uint4(q[0], q[1], q[2], q[3]);
    }
    // W21[j][k] = W2_j W1[j][k] in three bf16 pieces: B operand [n = 32*piece + k][j], K-major (chunks of 8 j)
    for (int idx = tid; idx < TK * 16; idx += TILE) {
        const int k = idx & 31, c = idx >> 5;
        uint32_t pc[3][4];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = 8 * c + e;
            uint32_t p0, p1, p2;
            umma::split_bf16x3(__ldg(w.W2 + j) * __ldg(w.W1 + j * TK + k), p0, p1, p2);
            if (e & 1) { pc[0][e >> 1] |= p0 << 16; pc[1][e >> 1] |= p1 << 16; pc[2][e >> 1] |= p2 << 16; }
            else { pc[0][e >> 1] = p0; pc[1][e >> 1] = p1; pc[2][e >> 1] = p2; }
        }
#pragma unroll
        for (int p = 0; p < 3; ++p)
            *reinterpret_cast<uint4 *>(smem + BS_W21 + c * (96 * 16) + (32 * p + k) * 16) = make_uint4(pc[p][0], pc[p][1], pc[p][2], pc[p][3]);
    }
    // the padding columns of the g*[u,1] tile never change: chunk 15 and the tails of the piece blocks are rewritten
    // with zeros by the per-tile stores below (k >= 33), chunk 15 here
    *reinterpret_cast<uint4 *>(smem + BS_GU + 15 * CHUNK_B + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) umma::mbar_init(bar, 1);
    if (warp == 0) umma::tmem_alloc(tslot, 256);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t t_du = *tslot, t_v = t_du + 128;
    const uint32_t trow = (uint32_t)(warp * 32) << 16;
    const uint32_t s_bits = umma::smem_u32(smem + BS_BITS), s_gu = umma::smem_u32(smem + BS_GU), s_w21 = umma::smem_u32(smem + BS_W21);
    constexpr uint32_t ID_DU = umma::idesc(UMMA_FMT_BF16, TILE, 96, false, false);
    constexpr uint32_t ID_V = umma::idesc(UMMA_FMT_BF16, TJ, 128, true, true);
    uint32_t phase = 0;
    // per-thread accumulators of the small gradients
    float a_b0[TH], a_ga[TH], a_be[TH], a_b2 = 0.f, a_rel[TH];
#pragma unroll
    for (int i = 0; i < TH; ++i) a_b0[i] = a_ga[i] = a_be[i] = a_rel[i] = 0.f;
    int cur_head = -1;                                           // warp-uniform head of a_rel (or -1)
    float w0a[TH / 2];
#pragma unroll
    for (int k = 0; k < TH / 2; ++k) w0a[k] = 0.f;
    const int i16 = lane & 15, half = lane >> 4;
    auto flush_rel = [&]() {                                     // called by whole warps
        if (cur_head >= 0) {
#pragma unroll
            for (int k = 0; k < TH; ++k) {
                const float s = warp_sumf(a_rel[k]);
                if (lane == 0 && s != 0.f) atomicAdd(grel + (size_t)cur_head * TH + k, s);
                a_rel[k] = 0.f;
            }
        }
    };
    for (long long tile = t0; tile < t1; ++tile) {
        const long long cell = tile * TILE + tid;
        const bool live = cell < C;
        float f[TH], nrm[TH];
        if (live) load_row16(F + cell * TH, f);
        else {
#pragma unroll
            for (int k = 0; k < TH; ++k) f[k] = 0.f;
        }
        const float g = live ? Gc[cell] : 0.f;
        const uint4 bw = live ? bits[cell] : make_uint4(0u, 0u, 0u, 0u);
        const int head = live ? slot_head[cell_key[cell] >> 5] : -1;
        const float rstd = cell_front<PRE>(sp, f, nrm);
        // ---- bit tile: 16 bytes -> 16 chunks of eight bf16 ----
        {
            const uint32_t bwv[4] = {bw.x, bw.y, bw.z, bw.w};
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const uint32_t b = (bwv[c >> 2] >> ((c & 3) * 8)) & 0xFFu;
                *reinterpret_cast<uint4 *>(smem + BS_BITS + c * CHUNK_B + tid * 16) = *reinterpret_cast<const uint4 *>(smem + BS_LUT + b * 16);
            }
        }
        // ---- g * [o, rel[head], 1] in three bf16 pieces ----
        {
            float r16[TH];
            if (live) load_row16(w.rel + (size_t)head * TH, r16);
            uint32_t pc[3][20];                                  // 40 bf16 per piece, packed
#pragma unroll
            for (int k = 0; k < 40; ++k) {
                float v = 0.f;
                if (k < TH) v = live ? g * fmaxf(fmaf(sp[SP_GA + k], nrm[k], sp[SP_BE + k]), 0.f) : 0.f;
                else if (k < TK) v = live ? g * r16[k - TH] : 0.f;
                else if (k == TK) v = g;
                uint32_t p0 = 0u, p1 = 0u, p2 = 0u;
                if (k <= TK) umma::split_bf16x3(v, p0, p1, p2);
                if (k & 1) { pc[0][k >> 1] |= p0 << 16; pc[1][k >> 1] |= p1 << 16; pc[2][k >> 1] |= p2 << 16; }
                else { pc[0][k >> 1] = p0; pc[1][k >> 1] = p1; pc[2][k >> 1] = p2; }
            }
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int c = 0; c < 5; ++c)
                    *reinterpret_cast<uint4 *>(smem + BS_GU + (5 * p + c) * CHUNK_B + tid * 16) =
                        make_uint4(pc[p][4 * c], pc[p][4 * c + 1], pc[p][4 * c + 2], pc[p][4 * c + 3]);
        }
        umma::fence_smem_to_async();
        umma::fence_before();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after();
#pragma unroll
            for (int k = 0; k < TJ / 16; ++k)                    // du: K = j, 16 per step = two chunks
                umma::mma_bf16(t_du, umma::desc(s_bits + k * 2 * CHUNK_B, CHUNK_B, 128),
                               umma::desc(s_w21 + k * 2 * (96 * 16), 96 * 16, 128), ID_DU, k != 0);
#pragma unroll
            for (int k = 0; k < TILE / 16; ++k)                  // V: K = cells, 16 rows per step
                umma::mma_bf16(t_v, umma::desc(s_bits + k * 256, 128, CHUNK_B), umma::desc(s_gu + k * 256, 128, CHUNK_B),
                               ID_V, tile != t0 || k != 0);
            umma::commit(bar);
        }
        umma::mbar_wait(bar, phase);
        phase ^= 1u;
        umma::fence_after();
        float du[TK];
        {
            float c0[32], c1[32], c2[32];
            umma::tmem_ld32(t_du + trow, c0);
            umma::tmem_ld32(t_du + trow + 32, c1);
            umma::tmem_ld32(t_du + trow + 64, c2);
#pragma unroll
            for (int k = 0; k < TK; ++k) du[k] = (c2[k] + c1[k]) + c0[k];
        }
        // relation half: summed per head relation (warp-uniform head: register accumulators; mixed warp: atomics)
        {
            const int h0 = __shfl_sync(FULL, head, 0);
            const bool uniform = __all_sync(FULL, head == h0 || head < 0);
            if (uniform) {
                if (h0 != cur_head && h0 >= 0) { flush_rel(); cur_head = h0; }
#pragma unroll
                for (int k = 0; k < TH; ++k) a_rel[k] = fmaf(g, du[TH + k], a_rel[k]);
            } else if (live) {
#pragma unroll
                for (int k = 0; k < TH; ++k) atomicAdd(grel + (size_t)head * TH + k, g * du[TH + k]);
            }
        }
        // ReLU + LayerNorm + Linear(H,H) backward
        float dn[TH], m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < TH; ++i) {
            const float ga = sp[SP_GA + i];
            const float pre = fmaf(ga, nrm[i], sp[SP_BE + i]);
            const float d_o = (live && pre > 0.f) ? g * du[i] : 0.f;
            a_ga[i] = fmaf(d_o, nrm[i], a_ga[i]);
            a_be[i] += d_o;
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
            a_b0[i] += dy;
            if (!PRE) {
                const float4 *wr = reinterpret_cast<const float4 *>(sp + SP_W0 + i * TH);
#pragma unroll
                for (int k4 = 0; k4 < TH / 4; ++k4) {
                    const float4 w4 = wr[k4];
                    dx[4 * k4] = fmaf(w4.x, dy, dx[4 * k4]); dx[4 * k4 + 1] = fmaf(w4.y, dy, dx[4 * k4 + 1]);
                    dx[4 * k4 + 2] = fmaf(w4.z, dy, dx[4 * k4 + 2]); dx[4 * k4 + 3] = fmaf(w4.w, dy, dx[4 * k4 + 3]);
                }
            }
        }
        a_b2 += g;
        if (live) {
            float4 *o1 = reinterpret_cast<float4 *>((PRE ? dY : dF) + cell * TH);
#pragma unroll
            for (int k4 = 0; k4 < TH / 4; ++k4)
                o1[k4] = PRE ? make_float4(dn[4 * k4], dn[4 * k4 + 1], dn[4 * k4 + 2], dn[4 * k4 + 3])
                             : make_float4(dx[4 * k4], dx[4 * k4 + 1], dx[4 * k4 + 2], dx[4 * k4 + 3]);
        }
        if (!PRE) {
            // dW0[i][k] += dy_i F_k: the warp stages 16 of its cells at a time, (unit i, half of k) lane pairs accumulate
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                __syncwarp();
                if ((lane >> 4) == pass) {
#pragma unroll
                    for (int k4 = 0; k4 < TH / 4; ++k4) {
                        *reinterpret_cast<float4 *>(stg + i16 * STG_STRIDE + 4 * k4) = make_float4(dn[4 * k4], dn[4 * k4 + 1], dn[4 * k4 + 2], dn[4 * k4 + 3]);
                        *reinterpret_cast<float4 *>(stg + (16 + i16) * STG_STRIDE + 4 * k4) = make_float4(f[4 * k4], f[4 * k4 + 1], f[4 * k4 + 2], f[4 * k4 + 3]);
                    }
                }
                __syncwarp();
#pragma unroll 8
                for (int c = 0; c < 16; ++c) {
                    const float dyi = stg[c * STG_STRIDE + i16];
                    const float4 *f4 = reinterpret_cast<const float4 *>(stg + (16 + c) * STG_STRIDE + half * 8);
                    const float4 fa = f4[0], fb = f4[1];
                    w0a[0] = fmaf(dyi, fa.x, w0a[0]); w0a[1] = fmaf(dyi, fa.y, w0a[1]); w0a[2] = fmaf(dyi, fa.z, w0a[2]); w0a[3] = fmaf(dyi, fa.w, w0a[3]);
                    w0a[4] = fmaf(dyi, fb.x, w0a[4]); w0a[5] = fmaf(dyi, fb.y, w0a[5]); w0a[6] = fmaf(dyi, fb.z, w0a[6]); w0a[7] = fmaf(dyi, fb.w, w0a[7]);
                }
            }
        }
    }
    flush_rel();
    // ---- once per kernel: V from TMEM (lane = hidden unit j), small gradients through shared memory ----
    {
        const int j = tid;
        float c0[32], c1[32];
        // columns: piece p at 40 p + k; add the three pieces of k = 0..31, then the g column (k = 32)
        float v[TK + 1];
        umma::tmem_ld32(t_v + trow, c0);                         // cols 0..31: piece 0, k 0..31
        umma::tmem_ld32(t_v + trow + 32, c1);                    // cols 32..63: piece 0 k 32 (col 32), piece 1 k 0..23 (cols 40..63)
#pragma unroll
        for (int k = 0; k < TK; ++k) v[k] = c0[k];
        v[TK] = c1[0];
#pragma unroll
        for (int k = 0; k < 24; ++k) v[k] += c1[8 + k];
        umma::tmem_ld32(t_v + trow + 64, c0);                    // cols 64..95: piece 1 k 24..32 (cols 64..72), piece 2 k 0..15 (cols 80..95)
#pragma unroll
        for (int k = 24; k <= TK; ++k) v[k] += c0[k - 24];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] += c0[16 + k];
        umma::tmem_ld32(t_v + trow + 96, c1);                    // cols 96..127: piece 2 k 16..32 (cols 96..112)
#pragma unroll
        for (int k = 16; k <= TK; ++k) v[k] += c1[k - 16];
#pragma unroll
        for (int k = 0; k <= TK; ++k) if (v[k] != 0.f) atomicAdd(Vg + (size_t)j * (TK + 1) + k, v[k]);
    }
#pragma unroll
    for (int i = 0; i < TH; ++i) {
        const float s0 = warp_sumf(a_b0[i]), s1 = warp_sumf(a_ga[i]), s2 = warp_sumf(a_be[i]);
        if (lane == 0) { atomicAdd(acc + TH * TH + i, s0); atomicAdd(acc + TH * TH + TH + i, s1); atomicAdd(acc + TH * TH + 2 * TH + i, s2); }
    }
    {
        const float s = warp_sumf(a_b2);
        if (lane == 0) atomicAdd(acc + TH * TH + 3 * TH, s);
    }
    if (!PRE) {
#pragma unroll
        for (int k = 0; k < TH / 2; ++k) atomicAdd(acc + i16 * TH + half * 8 + k, w0a[k]);
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_free(t_du, 256);
    if (!PRE)
        for (int i = tid; i < TH * TH; i += TILE) if (acc[i] != 0.f) atomicAdd(gW0 + i, acc[i]);
    if (tid < 3 * TH + 1) {
        const float v = acc[TH * TH + tid];
        if (v != 0.f) {
            float *dst = tid < TH ? gb0 + tid : tid < 2 * TH ? ggamma + (tid - TH) : tid < 3 * TH ? gbeta + (tid - 2 * TH) : gb2;
            atomicAdd(dst, v);
        }
    }
}

// dW1, db1, dW2 from V (one block, thread j); V[j][0..32) = sum bit_j g u_k, V[j][32] = sum bit_j g
__global__ void __launch_bounds__(TJ)
k_tail_finish(const float *__restrict__ W1, const float *__restrict__ b1, const float *__restrict__ W2,
              const float *__restrict__ Vg, float *__restrict__ gW1, float *__restrict__ gb1, float *__restrict__ gW2)
{
    const int j = threadIdx.x;
    const float w2 = W2[j], p = Vg[j * (TK + 1) + TK];
    float s = b1[j] * p;
    for (int k = 0; k < TK; ++k) {
        const float v = Vg[j * (TK + 1) + k];
        gW1[j * TK + k] += w2 * v;
        s = fmaf(W1[j * TK + k], v, s);
    }
    gb1[j] += w2 * p;
    gW2[j] += s;
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

/* floats of scratch rl_tail_backward needs (V [128][33]) */
int64_t rl_tail_scratch_floats(int32_t R) { (void)R; return (int64_t)TJ * (TK + 1); }

int rl_tail_forward(const rl_cells *c, const int32_t *slot_head, int32_t H, int32_t J, const float *F, const float *W0,
                    const float *b0, const float *gamma, const float *beta, const float *W1, const float *b1,
                    const float *W2, const float *b2, const float *rel_emb, float *zc, uint32_t *relu_bits,
                    int32_t front_done, void *stream)
{
    if (!c || !c->counters || !c->cell_key || !slot_head || !F || !W0 || !b0 || !gamma || !beta || !W1 || !b1 || !W2 || !b2 ||
        !rel_emb || !zc || !relu_bits)
        return rl_fail(RL_ERR_ARG, "rl_tail_forward: null argument");
    if (H != TH || J != TJ) return rl_fail(RL_ERR_ARG, "rl_tail_forward: built for hidden_dim 16 and a 128-wide score MLP");
    TailP w{W0, b0, gamma, beta, W1, b1, W2, b2, rel_emb};
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_tail_fwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_END);
        cudaFuncSetAttribute(k_tail_fwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_END);
        attr = true;
    }
    const long long want = ((long long)c->cap + TILE - 1) / TILE;
    const int grid = (int)(want < 3LL * sm_count() ? want : 3LL * sm_count());
    if (grid <= 0) return RL_OK;
    if (front_done) k_tail_fwd_tc<true><<<grid, TILE, FS_END, (cudaStream_t)stream>>>(c->counters, c->cap, F, c->cell_key, slot_head, w, zc,
                                                                                       reinterpret_cast<uint4 *>(relu_bits));
    else k_tail_fwd_tc<false><<<grid, TILE, FS_END, (cudaStream_t)stream>>>(c->counters, c->cap, F, c->cell_key, slot_head, w, zc,
                                                                             reinterpret_cast<uint4 *>(relu_bits));
    CHECK_LAUNCH("k_tail_fwd_tc");
    return RL_OK;
}

/* grads are ACCUMULATED; relu_bits come from rl_tail_forward of the same cells; scratch: rl_tail_scratch_floats(R) */
int rl_tail_backward(const rl_cells *c, const int32_t *slot_head, int32_t R, int32_t H, int32_t J, const float *F,
                     const float *W0, const float *b0, const float *gamma, const float *beta, const float *W1,
                     const float *b1, const float *W2, const float *b2, const float *rel_emb, const float *Gc,
                     const uint32_t *relu_bits, float *dF, float *dY, float *gW0, float *gb0,
                     float *ggamma, float *gbeta, float *gW1, float *gb1, float *gW2, float *gb2, float *grel,
                     float *scratch, int32_t front_done, void *stream)
{
    if (!c || !c->counters || !c->cell_key || !slot_head || !F || !W0 || !b0 || !gamma || !beta || !W1 || !b1 || !W2 || !b2 || !rel_emb ||
        !Gc || !relu_bits || (!front_done && (!dF || !gW0)) || (front_done && !dY) || !gb0 || !ggamma || !gbeta || !gW1 || !gb1 || !gW2 ||
        !gb2 || !grel || !scratch || R <= 0)
        return rl_fail(RL_ERR_ARG, "rl_tail_backward: null argument");
    if (H != TH || J != TJ) return rl_fail(RL_ERR_ARG, "rl_tail_backward: built for hidden_dim 16 and a 128-wide score MLP");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)rl_tail_scratch_floats(R) * sizeof(float), st);
    if (e != cudaSuccess) return rl_fail(RL_ERR_CUDA, "rl_tail_backward: memset", e);
    TailP w{W0, b0, gamma, beta, W1, b1, W2, b2, rel_emb};
    const uint4 *bits4 = reinterpret_cast<const uint4 *>(relu_bits);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_tail_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_END);
        cudaFuncSetAttribute(k_tail_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_END);
        attr = true;
    }
    const long long want = ((long long)c->cap + TILE - 1) / TILE;
    const int grid = (int)(want < 2LL * sm_count() ? want : 2LL * sm_count());
    if (grid <= 0) return RL_OK;
    if (front_done) k_tail_bwd_tc<true><<<grid, TILE, BS_END, st>>>(c->counters, c->cap, F, w, Gc, bits4, c->cell_key, slot_head, dF, dY, scratch,
                                                                     gW0, gb0, ggamma, gbeta, gb2, grel);
    else k_tail_bwd_tc<false><<<grid, TILE, BS_END, st>>>(c->counters, c->cap, F, w, Gc, bits4, c->cell_key, slot_head, dF, dY, scratch,
                                                           gW0, gb0, ggamma, gbeta, gb2, grel);
    CHECK_LAUNCH("k_tail_bwd_tc");
    k_tail_finish<<<1, TJ, 0, st>>>(W1, b1, W2, scratch, gW1, gb1, gW2);
    CHECK_LAUNCH("k_tail_finish");
    return RL_OK;
}

}  // extern "C"
