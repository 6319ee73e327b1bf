// Development tool (not on the product path): times the decode GEMM shapes through the C ABI with CUDA events,
// rotating over several weight copies so every launch streams from HBM, and dumps the per-role clock64 timeline
// of the first CTAs (gemm.cuh, B200_GEMM_TRACE).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr \
//        -DB200_GEMM_TRACE -o build/gemm_trace tools/gemm_trace.cu
#include "../lmdeploy_b200/csrc/capi_common.cu"
#include "../lmdeploy_b200/csrc/gemm.cu"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                                          \
    do {                                                                                                               \
        cudaError_t e_ = (x);                                                                                          \
        if (e_ != cudaSuccess) {                                                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);                            \
            exit(1);                                                                                                   \
        }                                                                                                              \
    } while (0)

__global__ void fill_kernel(uint32_t* p, size_t n, uint32_t seed)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)i * 2654435761u + seed;
        x ^= x >> 13;
        // keep fp16 halves small/finite when interpreted as params: mask exponent bits
        p[i] = x & 0x33ff33ffu;
    }
}

struct Shape {
    const char* name;
    int         K, N, epi, fmt;
};

int main(int argc, char** argv)
{
    const int m      = argc > 1 ? atoi(argv[1]) : 64;
    const int iters  = argc > 2 ? atoi(argv[2]) : 20;
    const int dump   = argc > 3 ? atoi(argv[3]) : 1;
    const int mt     = (m + 15) / 16 * 16;
    Shape     shapes[] = {{"w2", 14336, 4096, 0, 1}, {"qkv", 4096, 6144, 0, 1}, {"wo", 4096, 4096, 0, 1},
                          {"w1w3", 4096, 28672, 1, 1}, {"lm_head", 4096, 92544, 0, 0}};
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
#ifdef B200_GEMM_TRACE
    int* hang_host = nullptr;
    CK(cudaHostAlloc((void**)&hang_host, 64, cudaHostAllocMapped));
    memset(hang_host, 0, 64);
    {
        int* dptr = nullptr;
        CK(cudaHostGetDevicePointer((void**)&dptr, hang_host, 0));
        CK(cudaMemcpyToSymbol(b200::gemm::g_hang_host, &dptr, sizeof(dptr)));
    }
#define CKH(x)                                                                                                         \
    do {                                                                                                               \
        cudaError_t e_ = (x);                                                                                          \
        if (e_ != cudaSuccess) {                                                                                       \
            printf("CUDA error %s at line %d; hang record: set=%d tag=%d cta=(%d,%d) iter=%d thread=%d parity=%d\n",   \
                   cudaGetErrorString(e_), __LINE__, hang_host[0], hang_host[1], hang_host[2], hang_host[3],           \
                   hang_host[4], hang_host[5], hang_host[6]);                                                          \
            exit(1);                                                                                                   \
        }                                                                                                              \
    } while (0)
#else
#define CKH(x) CK(x)
#endif
    for (const Shape& s : shapes) {
        b200_linear_desc d{};
        d.input_dim = s.K, d.output_dim = s.N, d.weight_format = s.fmt, d.dtype = B200_DTYPE_F16, d.group_size = 128;
        const size_t wbytes = b200_linear_prepared_bytes(&d);
        const int    copies = (int)std::max<size_t>(2, (400u << 20) / wbytes);
        std::vector<void*> w(copies);
        for (int c = 0; c < copies; ++c) {
            CK(cudaMalloc(&w[c], wbytes));
            fill_kernel<<<1024, 256, 0, st>>>((uint32_t*)w[c], wbytes / 4, 17u * c + 1);
        }
        void*  xp;
        void*  y;
        void*  ws;
        size_t wsb = b200_linear_workspace_bytes(&d, m);
        CK(cudaMalloc(&xp, (size_t)s.K * mt * 2));
        CK(cudaMalloc(&y, (size_t)s.N * 128 * 2));
        CK(cudaMalloc(&ws, wsb));
        CK(cudaMemsetAsync(ws, 0, wsb, st));
        fill_kernel<<<256, 256, 0, st>>>((uint32_t*)xp, (size_t)s.K * mt / 2, 99);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        auto run = [&](int c) {
            d.prepared = w[c % copies];
            int rc = b200_linear_forward_packed(&d, xp, mt, m, y, s.epi ? s.N / 2 : s.N, s.epi, 0, ws, wsb, st);
            if (rc) {
                printf("forward failed: %s\n", b200_last_error());
                exit(1);
            }
        };
        for (int i = 0; i < 3; ++i) {
            run(i);
            CKH(cudaStreamSynchronize(st));
        }
#ifdef B200_GEMM_TRACE
        {
            int hang[8];
            CK(cudaMemcpyFromSymbol(hang, b200::gemm::g_hang, sizeof(hang)));
            if (hang[0]) {
                printf("%-8s m=%d HANG: wait tag %d (1 wempty/wprod, 2 xfull/mma, 3 aready/mma, 4 wfull/dequant, 5 dready, 6 xempty/xprod, 7 wfull/mma, 8 aempty/dequant) "
                       "cta (%d,%d) iter %d thread %d parity %d\n",
                       s.name, m, hang[1], hang[2], hang[3], hang[4], hang[5], hang[6]);
                memset(hang, 0, sizeof(hang));
                CK(cudaMemcpyToSymbol(b200::gemm::g_hang, hang, sizeof(hang)));
                continue;
            }
        }
#endif
        // steady-state time: the launches are captured into one CUDA graph (no CPU launch cost, warm i-cache),
        // exactly how the decode step runs them
        cudaGraph_t     graph;
        cudaGraphExec_t gexec;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
        for (int i = 0; i < iters; ++i) run(i + 3);
        CK(cudaStreamEndCapture(st, &graph));
        CK(cudaGraphInstantiate(&gexec, graph, 0));
        CK(cudaGraphLaunch(gexec, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventRecord(e0, st));
        CK(cudaGraphLaunch(gexec, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaGraphExecDestroy(gexec));
        CK(cudaGraphDestroy(graph));
        const double us = ms * 1e3 / iters;
        const b200::gemm::Plan pl = b200::gemm::make_plan(s.fmt, mt, s.N, s.K);
        printf("%-8s K=%5d N=%5d m=%d tiles=%d splits=%d nw=%d nx=%d na=%d  %8.2f us  %7.1f GB/s (weights %.1f MB)\n", s.name,
               s.K, s.N, m, pl.tiles, pl.splits, pl.nw, pl.nx, pl.na, us, wbytes / us * 1e-3, wbytes / 1e6);
#ifdef B200_GEMM_TRACE
        {
            // wall-clock anatomy of the last launches of the graph replay above (launch numbers are baked into the
            // captured kernel parameters; the replay overwrote the stamps of the capture-free warm-up launches)
            using namespace b200::gemm;
            static unsigned long long wall[kWallLaunches][kWallCtas][2];
            CK(cudaMemcpyFromSymbol(wall, g_wall, sizeof(wall)));
            const int ctas = std::min(kWallCtas, (int)((s.N / (128 * pl.tiles)) * pl.splits));
            // order the recorded launches by their first CTA start
            std::vector<std::pair<unsigned long long, int>> order;
            for (int l = 0; l < kWallLaunches; ++l) {
                unsigned long long b = ~0ull;
                for (int c = 0; c < ctas; ++c) b = std::min(b, wall[l][c][0]);
                if (b != ~0ull && b != 0) order.push_back({b, l});
            }
            std::sort(order.begin(), order.end());
            printf("  wall clock of %d consecutive launches (ns): [first CTA start .. last CTA start] -> [first CTA end .. last "
                   "CTA end], gap = this launch's first start - previous launch's last end\n", (int)order.size());
            unsigned long long prev_end = 0, t00 = order.empty() ? 0 : order[0].first;
            for (size_t i = 0; i < order.size(); ++i) {
                const int          l  = order[i].second;
                unsigned long long s0 = ~0ull, s1 = 0, e0 = ~0ull, e1 = 0;
                for (int c = 0; c < ctas; ++c) {
                    s0 = std::min(s0, wall[l][c][0]), s1 = std::max(s1, wall[l][c][0]);
                    e0 = std::min(e0, wall[l][c][1]), e1 = std::max(e1, wall[l][c][1]);
                }
                if (i >= order.size() - 6) {
                    printf("    launch %2zu: start %7llu .. %7llu  end %7llu .. %7llu   CTA lifetime %5llu   gap %6lld\n", i, s0 - t00,
                           s1 - t00, e0 - t00, e1 - t00, e1 - s0, prev_end ? (long long)(s0 - prev_end) : 0ll);
                }
                prev_end = e1;
            }
        }
        if (dump) {
            using namespace b200::gemm;
            static long long h[kTraceCtas][kTraceRoles][kTraceSlots];
            CK(cudaMemset(nullptr, 0, 0));
            long long* dptr;
            CK(cudaGetSymbolAddress((void**)&dptr, g_trace));
            CK(cudaMemsetAsync(dptr, 0, sizeof(h), st));
            run(1);
            CK(cudaStreamSynchronize(st));
            CK(cudaMemcpy(h, dptr, sizeof(h), cudaMemcpyDeviceToHost));
            for (int c = 0; c < 3; ++c) {
                const long long t0 = h[c][4][0];
                printf("  cta %d: start 0, tmem-alloc done %lld, dready %lld, epilogue done %lld\n", c, h[c][4][1] - t0,
                       h[c][4][2] - t0, h[c][4][3] - t0);
                printf("    after pdl_launch_dependents %lld, partials written+fence %lld, all partials arrived %lld, reduce done %lld, "
                       "kernel end %lld\n", h[c][7][4] - t0, h[c][7][0] - t0, h[c][7][1] - t0, h[c][7][2] - t0, h[c][7][3] - t0);
                printf("    %4s %9s %9s %9s %9s %9s %9s\n", "i", "w-issue", "x-issue", "w-full", "a-ready", "mma-issue", "commit");
                for (int i = 0; i < kTraceSlots && h[c][0][i]; ++i) {
                    printf("    %4d %9lld %9lld %9lld %9lld %9lld %9lld\n", i, h[c][0][i] - t0, h[c][6][i] - t0, h[c][1][i] - t0,
                           h[c][2][i] - t0, h[c][3][i] - t0, h[c][5][i] - t0);
                    if (i >= 40) break;
                }
            }
        }
#endif
        for (void* p : w) CK(cudaFree(p));
        CK(cudaFree(xp));
        CK(cudaFree(y));
        CK(cudaFree(ws));
    }
    return 0;
}
