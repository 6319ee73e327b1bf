// Development microbenchmark (not on the product path): issue rate / execution time of tcgen05.mma kind::f16 for
// the operand placements and shared-memory layouts gemm.cuh can choose from.  One thread per CTA issues R MMAs
// back to back into the same accumulator, commits, waits; reports SM cycles per MMA (max over CTAs).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o build/umma_rate tools/umma_rate.cu
#include "../lmdeploy_b200/csrc/umma.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace b200;

struct Variant {
    const char* name;
    int         ts;       // A from TMEM
    int         n;        // UMMA N
    int         layout;   // 0: no swizzle, lbo = rows*16 (k-chunk stride), sbo = 128
                          // 1: no swizzle, lbo = 128, sbo = 256 (core matrices of the two k-chunks adjacent)
                          // 2: SWIZZLE_128B K-major, sbo = 1024
    int         distinct; // 1: walk 8 different k-steps / 4 stages of descriptors like the real loop
    int         nacc;     // independent accumulators the MMAs rotate over
};

__device__ __forceinline__ uint64_t desc_for(uint32_t addr, int rows, int layout)
{
    if (layout == 0) {
        return umma::make_smem_desc(addr, rows * 16, 128);
    }
    if (layout == 1) {
        return umma::make_smem_desc(addr, 128, 256);
    }
    uint64_t d = umma::make_smem_desc(addr, 16, 1024);
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(128) rate_kernel(long long* out, int ts, int n, int layout, int distinct, int reps,
                                                   int tmem_cols, int nacc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t                       bar;
    __shared__ uint32_t                       tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) {
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
        umma::tmem_alloc(smem_u32(&tmem_base), tmem_cols);
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = tmem_base;
    const uint32_t a_col = (n * nacc <= 64) ? 64 : 256;
    if (ts) {
        uint32_t r[16];
        for (int c = 0; c < 16; ++c) r[c] = 0x3c003c00u;
        for (int c0 = 0; c0 < 128; c0 += 16) {
            umma::st_32x32b_x16(tb + ((uint32_t)(warp * 32) << 16) + a_col + c0, r);
        }
        umma::wait_st();
        umma::fence_before_sync();
    }
    __syncthreads();
    umma::fence_after_sync();
    long long t0 = 0, t1 = 0, t2 = 0;
    if (tid == 0) {
        const uint32_t idesc = umma::make_idesc_f16(128, n, false);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 48 * 1024;
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const int ks = distinct ? (r & 7) : 0;
            const int sg = distinct ? ((r >> 3) & 1) : 0;
            uint32_t  a_addr, b_addr;
            if (layout == 2) {
                a_addr = a0 + sg * 16384 + ks * 32;
                b_addr = b0 + sg * 16384 + ks * 32;
            }
            else if (layout == 1) {
                a_addr = a0 + sg * 16384 + ks * (128 * 32);
                b_addr = b0 + sg * 16384 + ks * (n * 32);
            }
            else {
                a_addr = a0 + sg * 16384 + ks * 2 * (128 * 16);
                b_addr = b0 + sg * 16384 + ks * 2 * (n * 16);
            }
            const uint64_t bd = desc_for(b_addr, n, layout);
            if (ts) {
                umma::mma_f16_ts(tb + (r % nacc) * n, tb + a_col + sg * 64 + ks * 8, bd, idesc, r >= nacc);
            }
            else {
                umma::mma_f16_ss(tb + (r % nacc) * n, desc_for(a_addr, 128, layout), bd, idesc, r >= nacc);
            }
        }
        t1 = clock64();
        umma::commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        t2 = clock64();
        out[blockIdx.x * 2]     = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        umma::tmem_dealloc(tb, tmem_cols);
    }
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    const int reps = 2048;
    Variant   vs[] = {
        {"TS  N=64  noswz 1 acc (gemm.cuh today)", 1, 64, 0, 1, 1},
        {"TS  N=64  noswz 2 acc", 1, 64, 0, 1, 2},
        {"TS  N=64  noswz 4 acc", 1, 64, 0, 1, 4},
        {"SS  N=64  noswz 1 acc (dense today)", 0, 64, 0, 1, 1},
        {"SS  N=64  noswz 4 acc", 0, 64, 0, 1, 4},
        {"TS  N=64  SW128 4 acc", 1, 64, 2, 1, 4},
        {"SS  N=64  SW128 4 acc", 0, 64, 2, 1, 4},
        {"TS  N=16  noswz 1 acc", 1, 16, 0, 1, 1},
        {"TS  N=16  noswz 4 acc", 1, 16, 0, 1, 4},
        {"TS  N=128 noswz 1 acc", 1, 128, 0, 1, 1},
        {"TS  N=128 noswz 2 acc", 1, 128, 0, 1, 2},
        {"SS  N=128 SW128 2 acc", 0, 128, 2, 1, 2},
        {"TS  N=128 SW128 2 acc", 1, 128, 2, 1, 2},
    };
    long long* d;
    cudaMalloc(&d, 1024 * 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
        const int grid = 148 * ctas_per_sm;
        for (const Variant& v : vs) {
            const int cols = (v.n * v.nacc <= 64) ? 256 : 512;
            if (ctas_per_sm == 2 && cols > 256) {
                continue;
            }
            cudaMemset(d, 0, 1024 * 16);
            rate_kernel<<<grid, 128, 96 * 1024>>>(d, v.ts, v.n, v.layout, v.distinct, reps, cols, v.nacc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("%-60s CUDA error %s\n", v.name, cudaGetErrorString(e));
                return 1;
            }
            std::vector<long long> h(grid * 2);
            cudaMemcpy(h.data(), d, grid * 16, cudaMemcpyDeviceToHost);
            long long mi = 0, mc = 0;
            for (int i = 0; i < grid; ++i) {
                mi = std::max(mi, h[2 * i]);
                mc = std::max(mc, h[2 * i + 1]);
            }
            printf("%d CTA/SM  %-60s issue %7.1f cyc/mma   complete %7.1f cyc/mma  (ideal %d)\n", ctas_per_sm, v.name,
                   (double)mi / reps, (double)mc / reps, 128 * v.n / 256);
        }
    }
    return 0;
}
