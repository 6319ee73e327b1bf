// Development tool (not on the product path): two questions the GEMM / attention producers hinge on.
//  (1) What does ISSUING a 1-D bulk copy (cp.async.bulk, SASS UBLKCP) cost the issuing thread, and does spreading the
//      issues over the lanes of a warp hide it?  (tools/gemm_trace timelines show ~300 cycles between back-to-back
//      issues from one lane; the attention producer got faster when ten lanes issued one copy each.)
//  (2) What per-SM and aggregate ingest rate do bulk copies reach as a function of copy size, copies in flight, and
//      whether the source is HBM (streamed once) or L2-resident (every SM re-reads the same 512 KB, as the activation
//      stream of the decode GEMMs does)?
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o build/tma_issue_rate tools/tma_issue_rate.cu
#include "../lmdeploy_b200/csrc/common.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace b200;

#define CK(x)                                                                                                          \
    do {                                                                                                               \
        cudaError_t e_ = (x);                                                                                          \
        if (e_ != cudaSuccess) {                                                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);                            \
            exit(1);                                                                                                   \
        }                                                                                                              \
    } while (0)

constexpr int kMaxDepth = 16;

// One warp per CTA, one CTA per SM.  `depth` slots of `bytes` each; `iters` rounds.  issue_lanes == 1: lane 0 issues
// every copy of a round back to back; issue_lanes == depth: lane i issues the copy of slot i.
// out[blockIdx.x] = {cycles spent issuing (sum over rounds, issuing lane max), total cycles}
// scatter != 0: the chunks of a CTA's region are visited in a pseudo-random order (multiplicative hash of the chunk
// index) instead of sequentially -- what a paged KV cache with a shuffled block table looks like to HBM.
__global__ void __launch_bounds__(32) ingest_kernel(const uint8_t* src, size_t src_stride_per_cta, size_t src_bytes_per_cta,
                                                    int bytes, int depth, int iters, int issue_lanes, long long* out,
                                                    int scatter)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t                       bar[kMaxDepth];
    const int                                 lane = threadIdx.x;
    if (lane == 0) {
        for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&bar[i]), 1);
        fence_mbar_init();
    }
    __syncwarp();
    const uint8_t* base  = src + (size_t)blockIdx.x * src_stride_per_cta;
    const size_t   nchunk = src_bytes_per_cta / bytes;
    auto           where = [&](size_t linear_off) -> size_t {
        if (!scatter) {
            return linear_off % src_bytes_per_cta;
        }
        const size_t c = (linear_off / bytes) % nchunk;
        return ((c * 2654435761ull + 12345u * blockIdx.x) % nchunk) * (size_t)bytes;
    };
    size_t         off   = 0;
    long long      t_iss = 0;
    const long long t0   = clock64();
    for (int r = 0; r < iters; ++r) {
        const uint32_t ph = r & 1;
        if (issue_lanes == 1) {
            if (lane == 0) {
                const long long a = clock64();
                for (int s = 0; s < depth; ++s) {
                    mbar_expect_tx(smem_u32(&bar[s]), bytes);
                    bulk_g2s(smem_u32(smem) + s * bytes, base + where(off + (size_t)s * bytes), bytes, smem_u32(&bar[s]));
                }
                t_iss += clock64() - a;
            }
        }
        else if (lane < depth) {
            const long long a = clock64();
            mbar_expect_tx(smem_u32(&bar[lane]), bytes);
            bulk_g2s(smem_u32(smem) + lane * bytes, base + where(off + (size_t)lane * bytes), bytes, smem_u32(&bar[lane]));
            t_iss += clock64() - a;
        }
        off += (size_t)depth * bytes;
        __syncwarp();
        for (int s = 0; s < depth; ++s) mbar_wait(smem_u32(&bar[s]), ph);
        __syncwarp();
    }
    const long long t1 = clock64();
    // issuing-lane maximum
    for (int o = 16; o; o >>= 1) t_iss = max(t_iss, __shfl_xor_sync(0xffffffffu, t_iss, o));
    if (lane == 0) {
        out[2 * blockIdx.x]     = t_iss;
        out[2 * blockIdx.x + 1] = t1 - t0;
    }
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const double ghz = prop.clockRate * 1e-6;
    printf("%s, %d SMs, %.2f GHz (nominal)\n", prop.name, sms, ghz);
    const size_t hbm_bytes = (size_t)2 << 30;  // streamed region: 2 GiB >> L2
    const size_t l2_bytes  = 512 << 10;        // shared hot region: 512 KB
    uint8_t *    d_hbm, *d_l2;
    long long*   d_out;
    CK(cudaMalloc(&d_hbm, hbm_bytes));
    CK(cudaMalloc(&d_l2, l2_bytes));
    CK(cudaMalloc(&d_out, sizeof(long long) * 2 * sms));
    CK(cudaMemset(d_hbm, 1, hbm_bytes));
    CK(cudaMemset(d_l2, 1, l2_bytes));
    CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    std::vector<long long> h(2 * sms);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("%-4s %7s %5s %6s | %10s %12s | %9s %9s\n", "src", "bytes", "depth", "lanes", "issue cyc", "/copy", "B/cyc/SM",
           "TB/s all");
    for (int l2 = 0; l2 < 2; ++l2) {
        for (int bytes : {2048, 4096, 8192, 16384, 32768}) {
            for (int depth : {1, 2, 4, 8}) {
                if (bytes * depth > 192 * 1024) {
                    continue;
                }
                for (int lanes : {1, depth}) {
                    if (lanes == 1 && depth == 1 && false) {
                        continue;
                    }
                    const size_t per_cta = l2 ? l2_bytes : hbm_bytes / sms / 1024 * 1024;
                    const int    iters   = (int)std::min<size_t>(400, std::max<size_t>(8, per_cta / ((size_t)bytes * depth)));
                    for (int rep = 0; rep < 2; ++rep) {  // first run warms the instruction cache (and L2 for the hot case)
                        CK(cudaEventRecord(e0));
                        ingest_kernel<<<sms, 32, bytes * depth>>>(l2 ? d_l2 : d_hbm, l2 ? 0 : per_cta, per_cta, bytes, depth,
                                                                 iters, lanes, d_out, 0);
                        CK(cudaEventRecord(e1));
                        CK(cudaDeviceSynchronize());
                    }
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost));
                    double iss = 0, tot = 0;
                    for (int i = 0; i < sms; ++i) iss += h[2 * i], tot += h[2 * i + 1];
                    iss /= sms, tot /= sms;
                    const double copies_per_lane = (double)iters * (lanes == 1 ? depth : 1);
                    const double total_bytes     = (double)sms * iters * depth * bytes;
                    printf("%-4s %7d %5d %6d | %10.0f %12.1f | %9.1f %9.2f\n", l2 ? "L2" : "HBM", bytes, depth, lanes, iss,
                           iss / copies_per_lane, (double)iters * depth * bytes / tot, total_bytes / (ms * 1e-3) / 1e12);
                    if (depth == 1) {
                        break;  // lanes == depth == 1: same configuration
                    }
                }
            }
        }
    }
    // (3) Paged-KV access pattern: every SM streams its own region of HBM in chunks of one (block, head) K or V tile
    // (8 KB for int8 KV, block_len 64), several copies in flight, chunks visited sequentially vs in shuffled order; with
    // 3 resident warps per SM (the attention kernel runs 3 CTAs per SM).  If "shuffled" sits near 0.75 of "sequential",
    // the decode-attention kernel is at the ceiling of this access pattern and only a larger contiguous unit
    // (block_len 128 -> 16 KB) can lift it.
    printf("\npaged-KV pattern (HBM, %d CTAs = 3 per SM, 8 copies in flight each, issued by 8 lanes):\n", 3 * sms);
    long long* d_out3;
    CK(cudaMalloc(&d_out3, sizeof(long long) * 2 * 3 * sms));
    for (int bytes : {4096, 8192, 16384}) {
        for (int scatter = 0; scatter < 2; ++scatter) {
            const int    depth   = 8;
            const size_t per_cta = hbm_bytes / (3 * sms) / 32768 * 32768;
            const int    iters   = (int)std::min<size_t>(2000, per_cta / ((size_t)bytes * depth));
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                ingest_kernel<<<3 * sms, 32, bytes * depth>>>(d_hbm, per_cta, per_cta, bytes, depth, iters, depth, d_out3,
                                                               scatter);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
            }
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double total_bytes = (double)3 * sms * iters * depth * bytes;
            printf("  chunk %6d B  %-10s  %7.2f TB/s  (%.1f MB in %.1f us)\n", bytes, scatter ? "shuffled" : "sequential",
                   total_bytes / (ms * 1e-3) / 1e12, total_bytes / 1e6, ms * 1e3);
        }
    }
    return 0;
}