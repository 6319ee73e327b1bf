// Standalone probe: validates the UMMA descriptor / TMEM layout assumptions used by lmdeploy_b200/csrc/gemm.cuh on a real
// sm_100a device.  Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/umma_probe tools/umma_probe.cu
// Prints one line per variant: max |D - expected|.  (Development tool; not part of the product path.)
#include "../lmdeploy_b200/csrc/umma.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace b200;

constexpr int M = 128, N = 16, K = 32;

__global__ void __launch_bounds__(128) probe(const half* A, const half* B, float* D, int variant)
{
    __shared__ __align__(128) uint8_t sA[M * K * 2];
    __shared__ __align__(128) uint8_t sB[N * K * 2 + 256];  // + room for the padded-stride variants
    __shared__ uint64_t               bar;
    __shared__ uint32_t               tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    // interleaved K-major layout: [k/8][row][8]
    for (int i = tid; i < M * K; i += 128) {
        int m = i / K, k = i % K;
        ((half*)sA)[(k / 8) * (M * 8) + m * 8 + k % 8] = A[m * K + k];
    }
    // variants 4/5: the stride between 8-row groups (SBO) is 144 B instead of 128 B -- core matrices stay 128 B
    // contiguous but successive ones start on different banks (would make the transposed V stores of the prefill
    // attention kernel conflict-free).  Does the hardware accept an SBO that is not a multiple of 128 B?
    const uint32_t b_sbo = variant >= 4 ? 144 : 128;
    const uint32_t b_lbo = (N / 8) * b_sbo;
    for (int i = tid; i < N * K; i += 128) {
        int n = i / K, k = i % K;
        *(half*)(sB + (k / 8) * b_lbo + (n / 8) * b_sbo + (n % 8) * 16 + (k % 8) * 2) = B[n * K + k];
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
        umma::tmem_alloc(smem_u32(&tmem_base), 64);
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb    = tmem_base;
    const uint32_t d_col = 0, a_col = 32;
    const bool     ts    = (variant & 2) != 0 || variant == 5;
    if (ts) {
        // A -> TMEM: lane = row, column c holds (k=2c, k=2c+1)
        uint32_t r[16];
        for (int c = 0; c < 16; ++c) {
            half2 h = __halves2half2(A[tid * K + 2 * c], A[tid * K + 2 * c + 1]);
            r[c]    = reinterpret_cast<uint32_t&>(h);
        }
        umma::st_32x32b_x16(tb + ((uint32_t)(warp * 32) << 16) + a_col, r);
        umma::wait_st();
        umma::fence_before_sync();
    }
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = umma::make_idesc_f16(M, N, false);
        const bool     swap  = variant < 4 && (variant & 1);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t a_addr = smem_u32(sA) + ks * 2 * (M * 16);
            const uint32_t b_addr = smem_u32(sB) + ks * 2 * b_lbo;
            uint64_t ad = swap ? umma::make_smem_desc(a_addr, 128, M * 16) : umma::make_smem_desc(a_addr, M * 16, 128);
            uint64_t bd = variant >= 4 ? umma::make_smem_desc(b_addr, b_lbo, b_sbo) :
                          swap         ? umma::make_smem_desc(b_addr, 128, N * 16) :
                                         umma::make_smem_desc(b_addr, N * 16, 128);
            if (ts) {
                umma::mma_f16_ts(tb + d_col, tb + a_col + ks * 8, bd, idesc, ks > 0);
            }
            else {
                umma::mma_f16_ss(tb + d_col, ad, bd, idesc, ks > 0);
            }
        }
        umma::commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    umma::fence_after_sync();
    uint32_t r[16];
    umma::ld_32x32b_x16(tb + ((uint32_t)(warp * 32) << 16) + d_col, r);
    umma::wait_ld();
    for (int j = 0; j < N; ++j) D[tid * N + j] = __uint_as_float(r[j]);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        umma::tmem_dealloc(tb, 64);
    }
}

int main()
{
    std::vector<half>  hA(M * K), hB(N * K);
    std::vector<float> ref(M * N), hD(M * N);
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) hA[m * K + k] = __float2half((float)((m * 3 + k * 5) % 7 - 3));
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) hB[n * K + k] = __float2half((float)((n * 2 + k) % 5 - 2));
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += __half2float(hA[m * K + k]) * __half2float(hB[n * K + k]);
            ref[m * N + n] = s;
        }
    half * dA, *dB;
    float* dD;
    cudaMalloc(&dA, M * K * 2);
    cudaMalloc(&dB, N * K * 2);
    cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
    const char* names[6] = {"SS lbo=kstride sbo=128", "SS lbo=128 sbo=kstride", "TS(A tmem) B lbo=kstride sbo=128",
                            "TS(A tmem) B lbo=128 sbo=kstride", "SS, B with sbo=144 (padded core-matrix stride)",
                            "TS(A tmem), B with sbo=144"};
    for (int v = 0; v < 6; ++v) {
        cudaMemset(dD, 0xff, M * N * 4);
        probe<<<1, 128>>>(dA, dB, dD, v);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("variant %d (%s): CUDA error %s\n", v, names[v], cudaGetErrorString(e));
            return 1;
        }
        cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        float mx = 0;
        int   bad = 0;
        for (int i = 0; i < M * N; ++i) {
            float d = fabsf(hD[i] - ref[i]);
            if (!(d <= 1e-3f)) ++bad;
            if (d > mx || d != d) mx = d;
        }
        printf("variant %d (%s): max_err=%g mismatches=%d/%d  D[0][0..3]=%g %g %g %g ref=%g %g %g %g  D[77][5]=%g ref=%g\n",
               v, names[v], mx, bad, M * N, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3], hD[77 * N + 5],
               ref[77 * N + 5]);
    }
    return 0;
}
