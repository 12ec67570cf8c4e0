// rnnlogic_b200 -- LSTM rule encoder of PredictorPlus (reference: src/predictors.py:141,201-208 --
// torch.nn.LSTM(hidden, hidden, num_layers, batch_first=True) over [head, body..., pad] embeddings, output
// of the last non-pad position).  The reference's FB15k-237 config is hidden_dim 16, 3 layers, sequences
// of <= 5 tokens, ~3e4 rules per step: far too small for cuDNN's RNN kernels (their element-wise backward
// alone takes milliseconds here).  Here H lanes own one rule (lane j = hidden unit j), weights sit in
// shared memory, the recurrence runs in registers, and only the sequential part is in these kernels:
//   forward  : out[n][H], the activations (i, f, g, o, c per layer and step) and, per layer and step, the pair
//              [layer input | previous hidden state] that the weight gradients contract with
//   backward : pre-activation gate gradients dG[l][n][t][4H] and the input gradient dX[n][t][H]
// The weight gradients are plain reductions over (rule, step) -- [dW_ih | dW_hh] = dG^T [In | Hprev],
// db = sum dG -- which the caller runs as ONE batched GEMM over the layers on the tensors written here.
// Positions behind a rule's last token do not influence its output (the recurrence is causal), so they are
// skipped: their activations / gradients stay at the caller's zero fill.
#include "rl_device.cuh"

#define RNN_THREADS 256

struct LstmW {
    const float *w_ih[RL_RNN_MAX_LAYERS], *w_hh[RL_RNN_MAX_LAYERS];   // [4H][H] row-major, gate order i, f, g, o
    const float *b_ih[RL_RNN_MAX_LAYERS], *b_hh[RL_RNN_MAX_LAYERS];   // [4H]
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// acts[n][L][T][5][H] = (i, f, g, o, c);  ih[L][n][T][2H] = [input of the layer | its hidden state of step t-1]
template <int H, int L>
__global__ void __launch_bounds__(RNN_THREADS)
k_lstm_fwd(int n, int T, const float *__restrict__ x, const int32_t *__restrict__ len, LstmW w,
           float *__restrict__ acts, float *__restrict__ ih, float *__restrict__ out)
{
    extern __shared__ float sm[];
    // transposed weights: wt_ih[l][k][4H], wt_hh[l][k][4H] (lane j reads consecutive words), bias[l][4H] = b_ih + b_hh
    float *wt_ih = sm, *wt_hh = sm + L * H * 4 * H, *bias = sm + 2 * L * H * 4 * H;
    for (int i = threadIdx.x; i < L * 4 * H * H; i += blockDim.x) {
        const int l = i / (4 * H * H), rem = i % (4 * H * H), row = rem / H, k = rem % H;
        wt_ih[(l * H + k) * 4 * H + row] = w.w_ih[l][rem];
        wt_hh[(l * H + k) * 4 * H + row] = w.w_hh[l][rem];
    }
    for (int i = threadIdx.x; i < L * 4 * H; i += blockDim.x) bias[i] = w.b_ih[i / (4 * H)][i % (4 * H)] + w.b_hh[i / (4 * H)][i % (4 * H)];
    __syncthreads();
    const int j = threadIdx.x % H;
    const int rule = (blockIdx.x * blockDim.x + threadIdx.x) / H;
    if (rule >= n) return;                                   // H divides 32: the lanes of a rule leave together
    const unsigned grp = H == 32 ? FULL : (0xffffu << ((threadIdx.x & 31) / H * H));
    const int steps = min(T, len[rule]);
    float h[L], c[L];
#pragma unroll
    for (int l = 0; l < L; ++l) { h[l] = 0.f; c[l] = 0.f; }
    for (int t = 0; t < steps; ++t) {
        float in = x[((size_t)rule * T + t) * H + j];
#pragma unroll
        for (int l = 0; l < L; ++l) {
            float *ip = ih + (((size_t)l * n + rule) * T + t) * 2 * H + j;
            ip[0] = in;
            ip[H] = h[l];
            float a[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) a[g] = bias[l * 4 * H + g * H + j];
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const float xk = __shfl_sync(grp, in, k, H), hk = __shfl_sync(grp, h[l], k, H);
                const float *wi = wt_ih + (l * H + k) * 4 * H + j, *wh = wt_hh + (l * H + k) * 4 * H + j;
#pragma unroll
                for (int g = 0; g < 4; ++g) a[g] = fmaf(wh[g * H], hk, fmaf(wi[g * H], xk, a[g]));
            }
            const float ig = sigmoidf_(a[0]), fg = sigmoidf_(a[1]), gg = tanhf(a[2]), og = sigmoidf_(a[3]);
            c[l] = fmaf(fg, c[l], ig * gg);
            h[l] = og * tanhf(c[l]);
            float *ap = acts + ((((size_t)rule * L + l) * T + t) * 5) * H + j;
            ap[0] = ig; ap[H] = fg; ap[2 * H] = gg; ap[3 * H] = og; ap[4 * H] = c[l];
            in = h[l];
        }
    }
    out[(size_t)rule * H + j] = h[L - 1];                      // output of the last non-pad position
}

// dG[l][n][t][4H] (pre-activation gate gradients), dX[n][t][H]; both zero-filled by the caller
template <int H, int L>
__global__ void __launch_bounds__(RNN_THREADS)
k_lstm_bwd(int n, int T, const int32_t *__restrict__ len, LstmW w, const float *__restrict__ acts,
           const float *__restrict__ dout, float *__restrict__ dG, float *__restrict__ dX)
{
    extern __shared__ float sm[];
    // original layout [l][4H][H]: lane k reads consecutive words of row (g, j)
    float *s_ih = sm, *s_hh = sm + L * 4 * H * H;
    for (int i = threadIdx.x; i < L * 4 * H * H; i += blockDim.x) {
        s_ih[i] = w.w_ih[i / (4 * H * H)][i % (4 * H * H)];
        s_hh[i] = w.w_hh[i / (4 * H * H)][i % (4 * H * H)];
    }
    __syncthreads();
    const int j = threadIdx.x % H;
    const int rule = (blockIdx.x * blockDim.x + threadIdx.x) / H;
    if (rule >= n) return;
    const unsigned grp = H == 32 ? FULL : (0xffffu << ((threadIdx.x & 31) / H * H));
    const int steps = min(T, len[rule]);
    float dh_rec[L], dc_rec[L];                                // gradients flowing back from step t+1
#pragma unroll
    for (int l = 0; l < L; ++l) { dh_rec[l] = 0.f; dc_rec[l] = 0.f; }
    for (int t = steps - 1; t >= 0; --t) {
        float from_above = (t == steps - 1) ? dout[(size_t)rule * H + j] : 0.f;     // d h[L-1][t]
#pragma unroll
        for (int l = L - 1; l >= 0; --l) {
            const float *ap = acts + ((((size_t)rule * L + l) * T + t) * 5) * H + j;
            const float ig = ap[0], fg = ap[H], gg = ap[2 * H], og = ap[3 * H], cc = ap[4 * H];
            const float cprev = t > 0 ? ap[4 * H - 5 * H] : 0.f;                     // c of step t-1
            const float dh = dh_rec[l] + from_above;
            const float tc = tanhf(cc);
            const float dc = fmaf(dh * og, 1.f - tc * tc, dc_rec[l]);
            float dg[4];
            dg[0] = dc * gg * ig * (1.f - ig);
            dg[1] = dc * cprev * fg * (1.f - fg);
            dg[2] = dc * ig * (1.f - gg * gg);
            dg[3] = dh * tc * og * (1.f - og);
            dc_rec[l] = dc * fg;
            float *gp = dG + (((size_t)l * n + rule) * T + t) * 4 * H + j;
#pragma unroll
            for (int g = 0; g < 4; ++g) gp[g * H] = dg[g];
            // lane k: d input[k] = sum_{g,j'} W_ih[g*H+j'][k] dgate[g][j'],  d h_prev[k] likewise with W_hh
            float dx = 0.f, dhp = 0.f;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
#pragma unroll
                for (int jj = 0; jj < H; ++jj) {
                    const float d = __shfl_sync(grp, dg[g], jj, H);
                    const int row = (l * 4 * H + g * H + jj) * H + j;
                    dx = fmaf(s_ih[row], d, dx);
                    dhp = fmaf(s_hh[row], d, dhp);
                }
            }
            dh_rec[l] = dhp;
            from_above = dx;                                   // gradient of layer l-1's output at this step
        }
        dX[((size_t)rule * T + t) * H + j] = from_above;
    }
}

// [dW_ih | dW_hh][l] = dG[l]^T ih[l]  ([4H] x [2H], reduced over the M = n*T (rule, step) rows) and db[l] = column
// sums of dG[l].  The shapes are far too skinny for a library GEMM (K = 1e5, 64 x 32 outputs): a block
// stages WG_TILE rows of both operands in shared memory, each thread owns one gate row and H*H/32 columns,
// and the block's partial result goes out with one atomic per output.  dW / db zero-filled by the caller.
#define WG_TILE 64
#define WG_ROWS_PER_BLOCK 1024
template <int H>
__global__ void __launch_bounds__(256)
k_lstm_wgrad(long long M, const float *__restrict__ dG, const float *__restrict__ ih, float *__restrict__ dW,
             float *__restrict__ db)
{
    constexpr int R = 4 * H, Cc = 2 * H, GROUPS = 256 / R, CPT = Cc / GROUPS;      // H=16: 64 rows, 4 groups, 8 columns each
    __shared__ float sG[WG_TILE][R], sI[WG_TILE][Cc];
    const int l = blockIdx.y;
    const float *G = dG + (size_t)l * M * R, *I = ih + (size_t)l * M * Cc;
    const int r = threadIdx.x % R, c0 = (threadIdx.x / R) * CPT;
    float acc[CPT], bsum = 0.f;
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = 0.f;
    const long long m0 = (long long)blockIdx.x * WG_ROWS_PER_BLOCK;
    const long long m1 = min(M, m0 + WG_ROWS_PER_BLOCK);
    for (long long mt = m0; mt < m1; mt += WG_TILE) {
        const int rows = (int)min((long long)WG_TILE, m1 - mt);
        for (int i = threadIdx.x; i < WG_TILE * R; i += 256) sG[i / R][i % R] = (i / R) < rows ? G[(size_t)mt * R + i] : 0.f;
        for (int i = threadIdx.x; i < WG_TILE * Cc; i += 256) sI[i / Cc][i % Cc] = (i / Cc) < rows ? I[(size_t)mt * Cc + i] : 0.f;
        __syncthreads();
#pragma unroll 4
        for (int m = 0; m < WG_TILE; ++m) {
            const float g = sG[m][r];
            bsum += g;
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[c] = fmaf(g, sI[m][c0 + c], acc[c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) atomicAdd(dW + ((size_t)l * R + r) * Cc + c0 + c, acc[c]);
    if (c0 == 0) atomicAdd(db + (size_t)l * R + r, bsum);
}

template <int H, int L>
static int lstm_launch(bool fwd, int n, int T, const float *x, const int32_t *len, const LstmW &w, float *acts, float *ih,
                       float *out, const float *dout, float *dG, float *dX, cudaStream_t st)
{
    const int rules_per_block = RNN_THREADS / H;
    const int grid = (n + rules_per_block - 1) / rules_per_block;
    if (fwd) {
        const size_t smem = (size_t)(2 * L * 4 * H * H + L * 4 * H) * sizeof(float);
        cudaFuncSetAttribute(k_lstm_fwd<H, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_lstm_fwd<H, L><<<grid, RNN_THREADS, smem, st>>>(n, T, x, len, w, acts, ih, out);
        CHECK_LAUNCH("k_lstm_fwd");
    } else {
        const size_t smem = (size_t)(2 * L * 4 * H * H) * sizeof(float);
        cudaFuncSetAttribute(k_lstm_bwd<H, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_lstm_bwd<H, L><<<grid, RNN_THREADS, smem, st>>>(n, T, len, w, acts, dout, dG, dX);
        CHECK_LAUNCH("k_lstm_bwd");
    }
    return RL_OK;
}

static int lstm_dispatch(bool fwd, int n, int T, int H, int L, const float *x, const int32_t *len, const float *const *weights,
                         float *acts, float *ih, float *out, const float *dout, float *dG, float *dX, void *stream)
{
    if (n <= 0) return RL_OK;
    if (!len || !weights || !acts || T < 1 || L < 1 || L > RL_RNN_MAX_LAYERS || (H != 16 && H != 32))
        return rl_fail(RL_ERR_ARG, "rl_lstm_encode: hidden_dim must be 16 or 32, 1..4 layers");
    LstmW w;
    for (int l = 0; l < RL_RNN_MAX_LAYERS; ++l) {
        const int s = l < L ? l : 0;
        w.w_ih[l] = weights[4 * s]; w.w_hh[l] = weights[4 * s + 1]; w.b_ih[l] = weights[4 * s + 2]; w.b_hh[l] = weights[4 * s + 3];
        if (!w.w_ih[l] || !w.w_hh[l] || !w.b_ih[l] || !w.b_hh[l]) return rl_fail(RL_ERR_ARG, "rl_lstm_encode: null weight");
    }
    cudaStream_t st = (cudaStream_t)stream;
#define RNN_CASE(HH, LL) if (H == HH && L == LL) return lstm_launch<HH, LL>(fwd, n, T, x, len, w, acts, ih, out, dout, dG, dX, st)
    RNN_CASE(16, 1); RNN_CASE(16, 2); RNN_CASE(16, 3); RNN_CASE(16, 4);
    RNN_CASE(32, 1); RNN_CASE(32, 2); RNN_CASE(32, 3); RNN_CASE(32, 4);
#undef RNN_CASE
    return rl_fail(RL_ERR_ARG, "rl_lstm_encode: unsupported shape");
}

extern "C" {

int rl_lstm_encode_forward(int32_t n, int32_t T, int32_t H, int32_t L, const float *x, const int32_t *len,
                           const float *const *weights, float *acts, float *ih, float *out, void *stream)
{
    if (n > 0 && (!x || !ih || !out)) return rl_fail(RL_ERR_ARG, "rl_lstm_encode_forward: null argument");
    return lstm_dispatch(true, n, T, H, L, x, len, weights, acts, ih, out, nullptr, nullptr, nullptr, stream);
}

int rl_lstm_encode_backward(int32_t n, int32_t T, int32_t H, int32_t L, const int32_t *len, const float *const *weights,
                            const float *acts, const float *dout, float *dG, float *dX, void *stream)
{
    if (n > 0 && (!dout || !dG || !dX)) return rl_fail(RL_ERR_ARG, "rl_lstm_encode_backward: null argument");
    return lstm_dispatch(false, n, T, H, L, nullptr, len, weights, const_cast<float *>(acts), nullptr, nullptr, dout, dG, dX, stream);
}

int rl_lstm_encode_wgrad(int64_t M, int32_t H, int32_t L, const float *dG, const float *ih, float *dW, float *db,
                         void *stream)
{
    if (M <= 0 || L <= 0) return RL_OK;
    if (!dG || !ih || !dW || !db || (H != 16 && H != 32)) return rl_fail(RL_ERR_ARG, "rl_lstm_encode_wgrad: bad argument");
    const dim3 grid((unsigned)((M + WG_ROWS_PER_BLOCK - 1) / WG_ROWS_PER_BLOCK), (unsigned)L);
    if (H == 16) k_lstm_wgrad<16><<<grid, 256, 0, (cudaStream_t)stream>>>(M, dG, ih, dW, db);
    else k_lstm_wgrad<32><<<grid, 256, 0, (cudaStream_t)stream>>>(M, dG, ih, dW, db);
    CHECK_LAUNCH("k_lstm_wgrad");
    return RL_OK;
}

}  // extern "C"
