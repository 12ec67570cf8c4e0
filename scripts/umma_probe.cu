// Development probe for csrc/rl_umma.cuh: runs single tcgen05.mma chains on hand-built un-swizzled shared-memory
// tiles in every operand role the tail kernels use and compares with a double-precision host product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I rnnlogic_b200/csrc -o gpurun_out/umma_probe scripts/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "rl_umma.cuh"

struct Job {
    int kind;            // 0 tf32, 1 bf16
    int a_mn, b_mn;      // operand majors
    int N;               // M = 128
    int ksteps;          // MMA instructions
    uint32_t a_bytes, b_bytes;          // tile images
    uint32_t a_lbo, a_sbo, a_step;      // descriptor fields / start-address advance per k step
    uint32_t b_lbo, b_sbo, b_step;
};

__global__ void __launch_bounds__(128) k_probe(Job j, const uint8_t *a_img, const uint8_t *b_img, float *D)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint8_t *sa = sm, *sb = sm + ((j.a_bytes + 127) & ~127u);
    for (uint32_t i = threadIdx.x * 16; i < j.a_bytes; i += 128 * 16) *reinterpret_cast<uint4 *>(sa + i) = *reinterpret_cast<const uint4 *>(a_img + i);
    for (uint32_t i = threadIdx.x * 16; i < j.b_bytes; i += 128 * 16) *reinterpret_cast<uint4 *>(sb + i) = *reinterpret_cast<const uint4 *>(b_img + i);
    if (threadIdx.x == 0) umma::mbar_init(&bar, 1);
    if (threadIdx.x < 32) umma::tmem_alloc(&tmem_base, 256);
    umma::fence_smem_to_async();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tb = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t id = umma::idesc(j.kind ? UMMA_FMT_BF16 : UMMA_FMT_TF32, 128, j.N, j.a_mn, j.b_mn);
        for (int k = 0; k < j.ksteps; ++k) {
            const uint64_t da = umma::desc(umma::smem_u32(sa) + k * j.a_step, j.a_lbo, j.a_sbo);
            const uint64_t db = umma::desc(umma::smem_u32(sb) + k * j.b_step, j.b_lbo, j.b_sbo);
            if (j.kind) umma::mma_bf16(tb, da, db, id, k > 0); else umma::mma_tf32(tb, da, db, id, k > 0);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after();
    const int warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < j.N; c0 += 16) {
        float v[16];
        umma::tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) D[(size_t)threadIdx.x * j.N + c0 + i] = v[i];
    }
    umma::fence_before();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_free(tb, 256);
}

static float bf16_round(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000u; memcpy(&x, &u, 4); return x; }
static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

// logical operand X[mn][k] -> tile image.  es = element bytes.  k_major: row = mn, chunk along k; else row = k, chunk along mn
static std::vector<uint8_t> image(const std::vector<float> &X, int MN, int K, int es, bool k_major)
{
    const int per = 16 / es;
    const int R = k_major ? MN : K;
    std::vector<uint8_t> img((size_t)MN * K * es, 0);
    for (int mn = 0; mn < MN; ++mn)
        for (int k = 0; k < K; ++k) {
            const int row = k_major ? mn : k, col = k_major ? k : mn;
            const size_t off = (size_t)(col / per) * R * 16 + (size_t)row * 16 + (size_t)(col % per) * es;
            const float x = X[(size_t)mn * K + k];
            uint32_t u; memcpy(&u, &x, 4);
            if (es == 4) memcpy(&img[off], &u, 4);
            else { const uint16_t h = (uint16_t)(u >> 16); memcpy(&img[off], &h, 2); }
        }
    return img;
}

static int run(const char *name, int kind, bool a_mn, bool b_mn, int N, int K)
{
    const int M = 128, es = kind ? 2 : 4, kper = kind ? 16 : 8;
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    for (auto &x : A) { x = (float)rand() / RAND_MAX - 0.5f; x = kind ? bf16_round(x) : tf32_trunc(x); }
    for (auto &x : B) { x = (float)rand() / RAND_MAX - 0.5f; x = kind ? bf16_round(x) : tf32_trunc(x); }
    auto ai = image(A, M, K, es, !a_mn), bi = image(B, N, K, es, !b_mn);
    Job j{};
    j.kind = kind; j.a_mn = a_mn; j.b_mn = b_mn; j.N = N; j.ksteps = K / kper;
    j.a_bytes = (uint32_t)ai.size(); j.b_bytes = (uint32_t)bi.size();
    // K-major: rows = MN -> LBO (next k chunk) = MN*16, SBO (next 8 rows) = 128, one k step = 2 chunks
    // MN-major: rows = K -> SBO (next mn chunk) = K*16, LBO (next 8 k rows) = 128, one k step = kper rows
    if (!a_mn) { j.a_lbo = M * 16; j.a_sbo = 128; j.a_step = 2 * M * 16; } else { j.a_sbo = K * 16; j.a_lbo = 128; j.a_step = kper * 16; }
    if (!b_mn) { j.b_lbo = N * 16; j.b_sbo = 128; j.b_step = 2 * N * 16; } else { j.b_sbo = K * 16; j.b_lbo = 128; j.b_step = kper * 16; }
    uint8_t *da, *db; float *dD;
    cudaMalloc(&da, ai.size()); cudaMalloc(&db, bi.size()); cudaMalloc(&dD, (size_t)M * N * 4);
    cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, (size_t)M * N * 4);
    const size_t smem = ((ai.size() + 127) & ~127ull) + bi.size() + 128;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<<<1, 128, smem>>>(j, da, db, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-40s CUDA error: %s\n", name, cudaGetErrorString(e)); return 2; }
    std::vector<float> D((size_t)M * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, scale = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
            worst = fmax(worst, fabs(s - (double)D[(size_t)m * N + n]));
            scale = fmax(scale, fabs(s));
        }
    printf("%-40s M=128 N=%d K=%d  max|err| %.3e (scale %.3e) %s\n", name, N, K, worst, scale, worst <= 1e-5 * scale ? "OK" : "MISMATCH");
    cudaFree(da); cudaFree(db); cudaFree(dD);
    return worst <= 1e-5 * scale ? 0 : 1;
}

int main()
{
    int bad = 0;
    bad |= run("tf32  A K-major  x B K-major", 0, false, false, 128, 32);
    bad |= run("bf16  A K-major  x B K-major", 1, false, false, 96, 128);
    bad |= run("bf16  A MN-major x B MN-major", 1, true, true, 128, 128);
    bad |= run("bf16  A MN-major x B K-major", 1, true, false, 64, 64);
    bad |= run("tf32  A K-major  x B K-major N=16", 0, false, false, 16, 64);
    bad |= run("tf32  A MN-major x B MN-major", 0, true, true, 128, 128);
    bad |= run("tf32  A K-major  x B K-major N=48", 0, false, false, 48, 64);
    bad |= run("tf32  A K-major  x B K-major N=64 K=48", 0, false, false, 64, 48);
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad;
}
