// Device-side primitives shared by the sm_100a kernels: mbarrier, bulk async copy (TMA 1-D),
// mma.sync fragments, sub-byte -> half/bf16 converters with the exact rounding contract of the
// reference (see DESIGN.md "numerics contract").
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

#define B200_DEVICE __device__ __forceinline__

// ------------------------------------------------------------------------------------------
// shared-memory address / mbarrier / bulk copy
// ------------------------------------------------------------------------------------------
B200_DEVICE uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

B200_DEVICE void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

B200_DEVICE void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

B200_DEVICE void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// generic <-> async proxy ordering for GLOBAL memory touched by bulk copies on both sides of a generic-proxy handshake
B200_DEVICE void fence_proxy_async_all()
{
    asm volatile("fence.proxy.async;" ::: "memory");
}

B200_DEVICE void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

B200_DEVICE void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

B200_DEVICE bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

B200_DEVICE void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
B200_DEVICE void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_smem),
                 "l"(src),
                 "r"(bytes),
                 "r"(bar)
                 : "memory");
}

// same with an L2 evict-first policy (data that is read exactly once per step)
B200_DEVICE void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(dst_smem),
        "l"(src),
        "r"(bytes),
        "r"(bar),
        "l"(policy)
        : "memory");
}

// 1-D bulk async copy shared -> global (bulk-group completion): commit, then wait until the writes are performed
B200_DEVICE void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
                 : "memory");
}
B200_DEVICE void bulk_commit_wait_all()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

B200_DEVICE uint64_t make_evict_first_policy()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// true in exactly one (converged) lane of the warp; keeps the surrounding code warp-uniform for ptxas
B200_DEVICE bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// Programmatic dependent launch (host_utils.h launch_k): allow the next kernel of the stream to start launching /
// wait until every kernel this one depends on has completed and its writes are visible.
B200_DEVICE void pdl_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
B200_DEVICE void pdl_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

B200_DEVICE void named_bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// non-blocking arrival on a named barrier other warps bar.sync on (`threads` = arrivals + waiters)
B200_DEVICE void named_bar_arrive(int id, int threads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ------------------------------------------------------------------------------------------
// shared loads
// ------------------------------------------------------------------------------------------
B200_DEVICE uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
B200_DEVICE uint2 lds64(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
B200_DEVICE uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// ------------------------------------------------------------------------------------------
// type traits for T in {half, nv_bfloat16}
// ------------------------------------------------------------------------------------------
template<class T>
struct TypeOps;

template<>
struct TypeOps<half> {
    using T2 = half2;
    static B200_DEVICE half  from_f32(float x) { return __float2half_rn(x); }
    static B200_DEVICE float to_f32(half x) { return __half2float(x); }
    // single-rounded T ops that ptxas may NOT contract
    static B200_DEVICE half mul(half a, half b) { return __hmul_rn(a, b); }
    static B200_DEVICE half add(half a, half b) { return __hadd_rn(a, b); }
    static B200_DEVICE half sub(half a, half b) { return __hsub_rn(a, b); }
    static B200_DEVICE half fma(half a, half b, half c) { return __hfma(a, b, c); }
    static B200_DEVICE half neg(half a) { return __hneg(a); }
    static B200_DEVICE half from_bits(uint16_t b) { return __ushort_as_half(b); }
    static B200_DEVICE uint16_t to_bits(half x) { return __half_as_ushort(x); }
    static B200_DEVICE uint32_t fma2(uint32_t a, uint32_t b, uint32_t c)
    {
        uint32_t d;
        asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
        return d;
    }
    static B200_DEVICE uint32_t pack2(float lo, float hi)
    {
        half2 h = __floats2half2_rn(lo, hi);
        return reinterpret_cast<uint32_t&>(h);
    }
    static B200_DEVICE uint32_t bcast2(half x)
    {
        half2 h = __half2half2(x);
        return reinterpret_cast<uint32_t&>(h);
    }
    // cvt.rni.sat.u8.f16
    static B200_DEVICE uint32_t round_u8(half x)
    {
        uint32_t y;
        asm("cvt.rni.sat.u8.f16 %0, %1;" : "=r"(y) : "h"(reinterpret_cast<uint16_t&>(x)));
        return y;
    }
    static B200_DEVICE void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
    {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    // two u8 (bytes sel lo / hi of `x`) -> T2 holding the exact integers
    // x bytes [b0 b1 b2 b3]; returns {(b0,b1), (b2,b3)} as exact halves (cvt_f16x4_u8, quantization.h:88-105)
    static B200_DEVICE void cvt_u8x4(uint32_t x, uint32_t& lo, uint32_t& hi)
    {
        lo = __byte_perm(x, 0x64000000u, 0x7170);
        hi = __byte_perm(x, 0x64000000u, 0x7372);
        asm("sub.rn.f16x2 %0, %0, %1;" : "+r"(lo) : "r"(0x64006400u));
        asm("sub.rn.f16x2 %0, %0, %1;" : "+r"(hi) : "r"(0x64006400u));
    }
    // u32 holding 8 nibbles n0..n7 -> 4 x T2: (n0,n4) (n1,n5) (n2,n6) (n3,n7) as exact halves
    // (cvt_f16x8_u4, quantization.h:491-524)
    static B200_DEVICE void cvt_u4x8(uint32_t x, uint32_t (&h)[4])
    {
        const uint32_t top = x >> 8;
        asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[0]) : "r"(x), "n"(0x000f000f), "n"(0x64006400));
        asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[1]) : "r"(x), "n"(0x00f000f0), "n"(0x54005400));
        asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[2]) : "r"(top), "n"(0x000f000f), "n"(0x64006400));
        asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[3]) : "r"(top), "n"(0x00f000f0), "n"(0x54005400));
        asm("sub.rn.f16x2 %0, %0, %1;" : "+r"(h[0]) : "r"(0x64006400u));
        asm("sub.rn.f16x2 %0, %0, %1;" : "+r"(h[1]) : "r"(0x54005400u));
        asm("sub.rn.f16x2 %0, %0, %1;" : "+r"(h[2]) : "r"(0x64006400u));
        asm("sub.rn.f16x2 %0, %0, %1;" : "+r"(h[3]) : "r"(0x54005400u));
    }
};

template<>
struct TypeOps<nv_bfloat16> {
    using T2 = nv_bfloat162;
    static B200_DEVICE nv_bfloat16 from_f32(float x) { return __float2bfloat16_rn(x); }
    static B200_DEVICE float       to_f32(nv_bfloat16 x) { return __bfloat162float(x); }
    static B200_DEVICE nv_bfloat16 mul(nv_bfloat16 a, nv_bfloat16 b) { return __hmul_rn(a, b); }
    static B200_DEVICE nv_bfloat16 add(nv_bfloat16 a, nv_bfloat16 b) { return __hadd_rn(a, b); }
    static B200_DEVICE nv_bfloat16 sub(nv_bfloat16 a, nv_bfloat16 b) { return __hsub_rn(a, b); }
    static B200_DEVICE nv_bfloat16 fma(nv_bfloat16 a, nv_bfloat16 b, nv_bfloat16 c) { return __hfma(a, b, c); }
    static B200_DEVICE nv_bfloat16 neg(nv_bfloat16 a) { return __hneg(a); }
    static B200_DEVICE nv_bfloat16 from_bits(uint16_t b) { return __ushort_as_bfloat16(b); }
    static B200_DEVICE uint16_t to_bits(nv_bfloat16 x) { return __bfloat16_as_ushort(x); }
    static B200_DEVICE uint32_t fma2(uint32_t a, uint32_t b, uint32_t c)
    {
        uint32_t d;
        asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
        return d;
    }
    static B200_DEVICE uint32_t pack2(float lo, float hi)
    {
        nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return reinterpret_cast<uint32_t&>(h);
    }
    static B200_DEVICE uint32_t bcast2(nv_bfloat16 x)
    {
        nv_bfloat162 h = __bfloat162bfloat162(x);
        return reinterpret_cast<uint32_t&>(h);
    }
    static B200_DEVICE uint32_t round_u8(nv_bfloat16 x)
    {
        uint32_t y;
        float    f = __bfloat162float(x);
        asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(y) : "f"(f));
        return y;
    }
    static B200_DEVICE void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
    {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    // integers 0..255 are exact in bf16 only through fp32 (cvt_bf16x4_u8, quantization.h:126-148)
    static B200_DEVICE void cvt_u8x4(uint32_t x, uint32_t& lo, uint32_t& hi)
    {
        float f0 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7650)) - 8388608.f;
        float f1 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7651)) - 8388608.f;
        float f2 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7652)) - 8388608.f;
        float f3 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7653)) - 8388608.f;
        lo       = pack2(f0, f1);
        hi       = pack2(f2, f3);
    }
    // nibble q -> bf16: 0x4300 | q = 128 + q (exact), minus 128
    static B200_DEVICE void cvt_u4x8(uint32_t x, uint32_t (&h)[4])
    {
        const uint32_t x1 = x >> 4, x2 = x >> 8, x3 = x >> 12;
        asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[0]) : "r"(x), "n"(0x000f000f), "n"(0x43004300));
        asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[1]) : "r"(x1), "n"(0x000f000f), "n"(0x43004300));
        asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[2]) : "r"(x2), "n"(0x000f000f), "n"(0x43004300));
        asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(h[3]) : "r"(x3), "n"(0x000f000f), "n"(0x43004300));
        const uint32_t neg128 = 0xC300C300u;  // -128 in both lanes
        const uint32_t one    = 0x3F803F80u;
        for (int i = 0; i < 4; ++i) h[i] = fma2(h[i], one, neg128);
    }
};

// 8x8 b16 transpose across the warp (movmatrix)
B200_DEVICE uint32_t transpose_8x8_b16(uint32_t x)
{
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

B200_DEVICE float exp2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

}  // namespace b200
