// tcgen05 / TMEM helpers of the rnnlogic_b200 tail kernels (sm_100a inline PTX, no library).
//
// Operand tiles are built IN the kernel (they are functions of per-cell data, not copies of global memory), so
// there is no TMA here: threads write the tile with 16-byte shared stores in the un-swizzled canonical UMMA layout,
// fence to the async proxy, and one thread issues the MMAs.
//
// Canonical un-swizzled tile of R "rows" x C "columns" of 16-byte chunks ("chunk" = 16 B = 4 tf32 or 8 bf16):
//       byte offset(row, chunk) = chunk * (R * 16) + row * 16
// i.e. a core matrix (8 rows x 16 B) is 128 contiguous bytes, the next 8 rows follow at +128 B and the next chunk at
// +R*16 B.  The SAME bytes serve two operand roles:
//   * K-major  (row = M/N index, chunk runs along K):   SBO = 128,    LBO = R*16
//   * MN-major (row = K index,   chunk runs along M/N): SBO = R*16,   LBO = 128
// (descriptor fields as in the PTX ISA "tcgen05 shared memory descriptor": start >> 4 | LBO >> 4 << 16 |
// SBO >> 4 << 32 | version 1 << 46 | layout type 0 (no swizzle) << 61).
#pragma once
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// instruction descriptor (kind::tf32 / kind::f16, fp32 accumulate): formats 1 = bf16, 2 = tf32
#define UMMA_FMT_BF16 1u
#define UMMA_FMT_TF32 2u
__host__ __device__ constexpr uint32_t idesc(uint32_t fmt, int M, int N, bool a_mn_major, bool b_mn_major)
{
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t id, bool accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(id), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t id, bool accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(id), "r"((uint32_t)accumulate) : "memory");
}
// arrive on an mbarrier when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared stores -> visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// bounded wait: a lost MMA completion traps (the launch fails) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (spins > (1u << 24)) __trap();
    }
}

// TMEM: one warp allocates `cols` (power of two >= 32) columns and publishes the base address in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t base, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(cols) : "memory");
}
// 32 consecutive fp32 columns of the calling thread's TMEM lane (warp w of the CTA owns lanes 32*(w%4) ..)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- number splitting ----
// x = hi + lo with hi a tf32 (low 13 mantissa bits zero); lo = x - hi is exact in fp32 and is cut to tf32 by the tensor
// core: A*B ~ Ahi*Bhi + Ahi*Blo + Alo*Bhi, relative error ~2^-21 per product ("3xTF32")
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo)
{
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}
// x = p0 + p1 + p2 with bf16 pieces (3 x 8 significand bits): exact partner of an operand that IS exact in bf16 (0/1 bits)
__device__ __forceinline__ void split_bf16x3(float x, uint32_t &p0, uint32_t &p1, uint32_t &p2)
{
    const uint32_t u0 = __float_as_uint(x) & 0xFFFF0000u;
    const float r1 = x - __uint_as_float(u0);
    const uint32_t u1 = __float_as_uint(r1) & 0xFFFF0000u;
    const float r2 = r1 - __uint_as_float(u1);
    // round the last piece to nearest (carry into the exponent is fine: still a bf16)
    const uint32_t u2 = (__float_as_uint(r2) + 0x8000u) & 0xFFFF0000u;
    p0 = u0 >> 16; p1 = u1 >> 16; p2 = u2 >> 16;
}

}  // namespace umma
