// Thin inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path: tcgen05.mma with shared-memory operands and TMEM
// accumulators (kind::tf32 here, used by the tile engine k2_conv3d_tc.cu; the z-march engine k2_conv3d_zm.cu issues
// kind::f16 with its own mma_f16 / idesc_f16 on the same descriptors), TMEM allocation / loads, mbarriers and the proxy
// fences between them.
//
// Shared-memory operand layout used throughout libmvsb200 (SWIZZLE_NONE, K-major "interleave" canonical form):
// a matrix of R rows x 8 tf32 (one MMA K step = 32 bytes) is stored as two K-chunk planes of R x 16 bytes,
//     byte(row r, k)  =  base + (k / 4) * LBO + r * 16 + (k % 4) * 4          LBO = plane stride
// so an 8-row x 16-byte core matrix is 128 contiguous bytes (SBO = 128) and -- the property the implicit-GEMM
// convolution relies on -- a window of 128 rows starting at ANY row s is again a canonical operand whose
// descriptor start address is simply base + 16*s.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mvsb200 {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// Bounded spin: a descriptor bug must trap, not hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
#pragma unroll 1   // keep the spin loop small: an unrolled copy per call site evicts the issue loop from the instruction cache
    for (uint32_t i = 0; i < (1u << 26); i++)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

// ---- fences --------------------------------------------------------------------------------------------------
// generic-proxy st.shared -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------------------
// One full warp; ncols a power of two in [32, 512]; the base address is written to *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 8 / 16 consecutive 32-bit columns: thread t of the warp receives lane (taddr.lane + t).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v)
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v)
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors -----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE, K-major (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (stride between the two 16-byte K chunks)
//   [32,46) stride byte offset >> 4 (stride between 8-row groups) | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// Instruction descriptor for kind::tf32, fp32 accumulate, A and B both K-major (cute::UMMA::InstrDescriptor):
//   [4,6) D format 1 = f32 | [7,10) A format 2 = tf32 | [10,13) B format 2 = tf32 | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// mbarrier arrives once every MMA issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Split x into a tf32-exact high part and a tf32-exact remainder (3xTF32 error-compensated products).  The tensor
// core TRUNCATES fp32 operands to tf32 (measured, tests/probe/umma_probe.cu); rounding both parts to nearest here
// keeps the residual error unbiased.
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo)
{
    hi = round_tf32(x);
    lo = round_tf32(x - hi);
}

}  // namespace umma
}  // namespace mvsb200
