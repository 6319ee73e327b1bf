// Development tool (not on the product path): measures the launch-plan space of the decode GEMMs -- n-tiles per CTA,
// K splits, ring depths -- per shape, graph-timed exactly like tools/gemm_trace.cu, and checks every plan's output
// against the default plan's on the same inputs.  One GPU call answers "is make_plan's heuristic leaving time on the
// table for this shape?" (the reference does the same search at run time: gemm/tuner, dispatch_cache.cu).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr \
//        -DB200_GEMM_TOOL -o build/gemm_sweep tools/gemm_sweep.cu
//   build/gemm_sweep [model: 7b|20b|8b|70b-tp8|all] [m=64] [iters=20] [tiles mask=3]
#include "../lmdeploy_b200/csrc/capi_common.cu"
#include "../lmdeploy_b200/csrc/gemm.cu"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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
        p[i] = x & 0x33ff33ffu;  // small finite fp16 halves when read as parameters
    }
}

__global__ void max_diff_kernel(const half* a, const half* b, size_t n, float* out)
{
    float m = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float d = fabsf(__half2float(a[i]) - __half2float(b[i]));
        m             = d == d ? fmaxf(m, d) : INFINITY;
    }
    atomicMax((int*)out, __float_as_int(m));  // non-negative floats order like ints
}

struct Shape {
    const char* model;
    const char* name;
    int         K, N, epi, fmt;
};

static const Shape kShapes[] = {
    {"7b", "qkv", 4096, 6144, 0, 1},          {"7b", "wo", 4096, 4096, 0, 1},
    {"7b", "w1w3", 4096, 28672, 1, 1},        {"7b", "w2", 14336, 4096, 0, 1},
    {"20b", "qkv", 6144, 8192, 0, 1},         {"20b", "wo", 6144, 6144, 0, 1},
    {"20b", "w1w3", 6144, 32768, 1, 1},       {"20b", "w2", 16384, 6144, 0, 1},
    {"8b", "qkv", 4096, 6144, 0, 0},          {"8b", "wo", 4096, 4096, 0, 0},
    {"8b", "w1w3", 4096, 28672, 1, 0},        {"8b", "w2", 14336, 4096, 0, 0},
    {"70b-tp8", "qkv", 8192, 1280, 0, 3},     {"70b-tp8", "wo", 1024, 8192, 0, 3},
    {"70b-tp8", "w1w3", 8192, 7168, 1, 3},    {"70b-tp8", "w2", 3584, 8192, 0, 3},
    // per-rank shapes of InternLM2-7B under tensor parallelism (qkv / w1w3 column-split, wo / w2 row-split)
    {"7b-tp2", "qkv", 4096, 3072, 0, 1},      {"7b-tp2", "wo", 2048, 4096, 0, 1},
    {"7b-tp2", "w1w3", 4096, 14336, 1, 1},    {"7b-tp2", "w2", 7168, 4096, 0, 1},
    {"7b-tp4", "qkv", 4096, 1536, 0, 1},      {"7b-tp4", "wo", 1024, 4096, 0, 1},
    {"7b-tp4", "w1w3", 4096, 7168, 1, 1},     {"7b-tp4", "w2", 3584, 4096, 0, 1},
    {"7b-tp8", "qkv", 4096, 768, 0, 1},       {"7b-tp8", "wo", 512, 4096, 0, 1},
    {"7b-tp8", "w1w3", 4096, 3584, 1, 1},     {"7b-tp8", "w2", 1792, 4096, 0, 1},
    {"20b-tp2", "qkv", 6144, 4096, 0, 1},     {"20b-tp2", "wo", 3072, 6144, 0, 1},
    {"20b-tp2", "w1w3", 6144, 16384, 1, 1},   {"20b-tp2", "w2", 8192, 6144, 0, 1},
    {"20b-tp4", "qkv", 6144, 2048, 0, 1},     {"20b-tp4", "wo", 1536, 6144, 0, 1},
    {"20b-tp4", "w1w3", 6144, 8192, 1, 1},    {"20b-tp4", "w2", 4096, 6144, 0, 1},
};

int main(int argc, char** argv)
{
    const std::string model = argc > 1 ? argv[1] : "7b";
    const int         m     = argc > 2 ? atoi(argv[2]) : 64;
    const int         iters = argc > 3 ? atoi(argv[3]) : 20;
    const int         tmask = argc > 4 ? atoi(argv[4]) : 3;  // which tiles-per-CTA values to sweep: bit0 1, bit1 2, bit2 4
                                                             // (4 = the tool-only kernel variant: run it in its own call)
    const int         mt    = (m + 15) / 16 * 16;
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    float* d_diff;
    CK(cudaMalloc(&d_diff, 4));
    for (const Shape& s : kShapes) {
        if (model != "all" && model != s.model) {
            continue;
        }
        b200_linear_desc d{};
        d.input_dim = s.K, d.output_dim = s.N, d.weight_format = s.fmt, d.dtype = B200_DTYPE_F16, d.group_size = 128;
        const size_t wbytes = b200_linear_prepared_bytes(&d);
        if (!wbytes) {
            printf("%s %s: %s\n", s.model, s.name, b200_last_error());
            continue;
        }
        const int          copies = (int)std::max<size_t>(2, (400u << 20) / wbytes);
        std::vector<void*> w(copies);
        for (int c = 0; c < copies; ++c) {
            CK(cudaMalloc(&w[c], wbytes));
            fill_kernel<<<1024, 256, 0, st>>>((uint32_t*)w[c], wbytes / 4, 17u * c + 1);
        }
        const size_t out_elems = (size_t)m * (s.epi ? s.N / 2 : s.N);
        void *       xp, *y, *y_ref, *ws;
        const size_t wsb = b200_linear_workspace_bytes(&d, m);
        CK(cudaMalloc(&xp, (size_t)s.K * mt * 2));
        CK(cudaMalloc(&y, out_elems * 2));
        CK(cudaMalloc(&y_ref, out_elems * 2));
        CK(cudaMalloc(&ws, wsb));
        CK(cudaMemsetAsync(ws, 0, wsb, st));
        fill_kernel<<<256, 256, 0, st>>>((uint32_t*)xp, (size_t)s.K * mt / 2, 99);
        auto run = [&](int c, void* out) {
            d.prepared = w[c % copies];
            const int rc = b200_linear_forward_packed(&d, xp, mt, m, out, s.epi ? s.N / 2 : s.N, s.epi, 0, ws, wsb, st);
            if (rc) {
                printf("forward failed: %s\n", b200_last_error());
                exit(1);
            }
        };
        auto time_plan = [&]() -> double {
            for (int i = 0; i < 2; ++i) run(i, y);
            CK(cudaStreamSynchronize(st));
            cudaGraph_t     graph;
            cudaGraphExec_t gexec;
            cudaEvent_t     e0, e1;
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
            for (int i = 0; i < iters; ++i) run(i + 3, y);
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
            CK(cudaEventDestroy(e0));
            CK(cudaEventDestroy(e1));
            return ms * 1e3 / iters;
        };
        using b200::gemm::g_plan_override;
        using b200::gemm::make_plan;
        using b200::gemm::Plan;
        g_plan_override   = {};
        const Plan   def  = make_plan(s.fmt, mt, s.N, s.K);
        const double t_def = time_plan();
        run(0, y_ref);
        CK(cudaStreamSynchronize(st));
        printf("%-8s %-5s K=%5d N=%5d m=%d  default: tiles=%d splits=%d nw=%d nx=%d  %7.2f us  %7.1f GB/s\n", s.model, s.name,
               s.K, s.N, m, def.tiles, def.splits, def.nw, def.nx, t_def, wbytes / t_def * 1e-3);
        // producer experiments on the default plan (gemm.cuh, Params::variant)
        for (int variant : {1, 2, 4, 5, 6}) {
            if (s.fmt == 0 || !(tmask & 3)) {
                break;
            }
            if ((variant & 2) && b200::gemm::block_bytes_of(s.fmt) % 32) {
                continue;  // half of the block must still be a multiple of the 16-byte bulk-copy granule
            }
            g_plan_override         = {};
            g_plan_override.variant = variant;
            const double t = time_plan();
            CK(cudaMemsetAsync(d_diff, 0, 4, st));
            run(0, y);
            max_diff_kernel<<<256, 256, 0, st>>>((const half*)y, (const half*)y_ref, out_elems, d_diff);
            float diff;
            CK(cudaMemcpyAsync(&diff, d_diff, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            printf("    producer variant %d (W: %s; X: %s): %7.2f us  (%+5.1f%%)  max|y - y_default| = %g\n", variant,
                   (variant & 3) == 0 ? "one lane issues every copy" : (variant & 3) == 1 ? "one lane per tile block" :
                                                                        "two half-block copies per tile, one lane each",
                   (variant & 4) ? "four quarter copies, four lanes" : "one copy", t, (t / t_def - 1) * 100, diff);
        }
        g_plan_override = {};
        double best = t_def;
        char   best_desc[128] = "default";
        char   best_plan[128] = "";
        const int kgs = s.K / 128;
        for (int tiles : {1, 2, 4}) {
            // 4 tiles per CTA: tool-only kernel variant (u4, <= 64 tokens); dense weights: one tile
            if (!(tmask & tiles) || s.N % (128 * tiles) || (s.fmt == 0 && tiles > 1) || (tiles == 4 && (s.fmt != 1 || mt > 64))) {
                continue;
            }
            const int nts = s.N / (128 * tiles);
            for (int splits : {1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 16}) {
                if (nts * splits > 148 || kgs / splits < 1) {
                    continue;  // cooperative launch needs one wave (uneven K ranges are fine: kg0 = kgs * sp / splits)
                }
                for (int depth = 0; depth < 3; ++depth) {
                    g_plan_override = {tiles, splits, 0, 0, 0};
                    const Plan base = make_plan(s.fmt, mt, s.N, s.K);
                    if (depth == 1) {  // deeper X ring, shallower W ring
                        g_plan_override = {tiles, splits, std::max(3, base.nw - 2), std::min(6, base.nx + 1), 0};
                    }
                    else if (depth == 2) {  // shallower X ring, deeper W ring
                        g_plan_override = {tiles, splits, std::min(8, base.nw + 1), std::max(2, base.nx - 1), 0};
                    }
                    const Plan pl = make_plan(s.fmt, mt, s.N, s.K);
                    if (pl.smem > 226 * 1024 || (depth && pl.nw == base.nw && pl.nx == base.nx)) {
                        continue;
                    }
                    const double t = time_plan();
                    CK(cudaMemsetAsync(d_diff, 0, 4, st));
                    run(0, y);
                    max_diff_kernel<<<256, 256, 0, st>>>((const half*)y, (const half*)y_ref, out_elems, d_diff);
                    float diff;
                    CK(cudaMemcpyAsync(&diff, d_diff, 4, cudaMemcpyDeviceToHost, st));
                    CK(cudaStreamSynchronize(st));
                    printf("    tiles=%d splits=%2d nw=%d nx=%d ctas=%3d  %7.2f us  (%+5.1f%%)  max|y - y_default| = %g\n", pl.tiles,
                           pl.splits, pl.nw, pl.nx, nts * splits, t, (t / t_def - 1) * 100, diff);
                    if (t < best) {
                        best = t;
                        snprintf(best_desc, sizeof(best_desc), "tiles=%d splits=%d nw=%d nx=%d", pl.tiles, pl.splits, pl.nw, pl.nx);
                        // only plans whose output matched the default's to fp16 rounding noise are worth recording
                        if (diff <= 1e-2f) {
                            snprintf(best_plan, sizeof(best_plan), "PLAN %d %d %d %d %d %d %d %d", s.fmt, mt, s.N, s.K, pl.tiles,
                                     pl.splits, pl.nw, pl.nx);
                        }
                    }
                }
            }
        }
        printf("  -> best %s: %.2f us (default %.2f us, %.1f%% faster), %.1f GB/s of weights\n", best_desc, best, t_def,
               (1 - best / t_def) * 100, wbytes / best * 1e-3);
        if (best_plan[0] && best < 0.98 * t_def) {
            printf("%s\n", best_plan);  // grep '^PLAN' > plans.txt; B200_GEMM_PLANS=plans.txt or b200_linear_load_plans()
        }
        g_plan_override = {};
        for (void* p : w) CK(cudaFree(p));
        CK(cudaFree(xp));
        CK(cudaFree(y));
        CK(cudaFree(y_ref));
        CK(cudaFree(ws));
    }
    return 0;
}
