// tcgen05 (UMMA) / TMEM primitives for sm_100a, written as inline PTX.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptors"; cross-checked against the
// field definitions in CUTLASS 4.x cute/arch/mma_sm100_desc.hpp (header tree vendored in the image).
#pragma once

#include "common.cuh"

namespace b200 {
namespace umma {

// ---- shared-memory matrix descriptor (K-major, no swizzle "interleave" canonical layout) ----
// In 16-byte units the operand is ((8,n),2):((1,SBO),LBO): a core matrix is 8 rows x 16 bytes,
// contiguous (128 B); SBO = byte stride between 8-row groups along M/N, LBO = byte stride between
// the two 16-byte K chunks of one K=32-byte MMA step.
B200_DEVICE uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version = 1 (Blackwell)
    // base_offset = 0, lbo_mode = 0, layout_type = 0 (SWIZZLE_NONE)
    return d;
}

// ---- instruction descriptor, kind::f16 (A,B in {f16,bf16}, fp32 accumulate), both K-major ----
B200_DEVICE constexpr uint32_t make_idesc_f16(int m, int n, bool bf16)
{
    uint32_t d = 0;
    d |= 1u << 4;                      // c_format = F32
    d |= (bf16 ? 1u : 0u) << 7;        // a_format
    d |= (bf16 ? 1u : 0u) << 10;       // b_format
    d |= (uint32_t)(n >> 3) << 17;     // n_dim
    d |= (uint32_t)(m >> 4) << 24;     // m_dim
    return d;
}

// kind::f8f6f4: a/b formats 0 = e4m3, 1 = e5m2
B200_DEVICE constexpr uint32_t make_idesc_f8(int m, int n)
{
    uint32_t d = 0;
    d |= 1u << 4;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}

B200_DEVICE void tmem_alloc(uint32_t smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

B200_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

B200_DEVICE void fence_before_sync()
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
B200_DEVICE void fence_after_sync()
{
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]
B200_DEVICE void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem),
        "l"(b_desc),
        "r"(idesc),
        "r"(accumulate)
        : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]
B200_DEVICE void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc),
        "l"(b_desc),
        "r"(idesc),
        "r"(accumulate)
        : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when complete
B200_DEVICE void commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

B200_DEVICE void wait_st()
{
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
B200_DEVICE void wait_ld()
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 lanes x 32-bit, 16 consecutive columns: thread i of the warp <-> TMEM lane (lane_base + i)
B200_DEVICE void st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
            taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

B200_DEVICE void ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

B200_DEVICE void ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

}  // namespace umma
}  // namespace b200
