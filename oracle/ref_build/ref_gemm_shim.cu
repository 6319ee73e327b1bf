// TEST INFRASTRUCTURE, not product code.
// Thin C-ABI shim around the REFERENCE's own weight-only-quant GEMM (compiled in place from /root/reference by
// oracle/ref_build/Makefile into oracle/_ref/libtm_ref_gemm.so; no reference source is copied into this repo), so that
// bench / tools can time the unmodified TurboMind W4A16 path on the same B200 next to ours ("the kernel to
// beat").  The reference's sm80 s16816 kernels have empty bodies when compiled natively for sm_100 (their Arch tag is
// [800, 900), gemm_universal.h:174-175); the Makefile therefore compiles the kernel translation units from compute_80
// PTX to sm_100 SASS -- the same code the reference executes on Blackwell through the PTX JIT.
//
// The two entry points call, in the reference's own order, what
//   LinearWeight::prepare      src/turbomind/models/linear_weight.cc:101-324   (u4 branch: GetConverters, extend_to_u16,
//                                                                               Convert, fuse_scales_and_zeros, Convert)
//   LlamaLinear::Impl::Forward src/turbomind/models/llama/LlamaLinear.cu:140-216 (descriptors + gemm::Gemm::Run)
// call, on raw device buffers instead of the reference's Tensor / Module classes (which would pull its whole model
// runtime into the build).
#include "src/turbomind/core/logger.h"
#include "src/turbomind/kernels/core/data_type.h"
#include "src/turbomind/kernels/gemm/cast.h"
#include "src/turbomind/kernels/gemm/convert.h"
#include "src/turbomind/kernels/gemm/gemm.h"
#include "src/turbomind/kernels/gemm/types.h"
#include "src/turbomind/utils/cuda_utils.h"

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>

namespace turbomind::core {
Logger::Logger() = default;
Logger& Logger::Instance()
{
    static thread_local Logger* inst = new Logger();
    return *inst;
}
void Logger::LogFatalImpl(const char* file, int line, std::string message)
{
    fprintf(stderr, "[ref-gemm-shim fatal] %s:%d %s\n", file ? file : "?", line, message.c_str());
    abort();
}
void Logger::LogFatalTraced(const char* file, int line, std::string message)
{
    LogFatalImpl(file, line, std::move(message));
}
void Logger::Enqueue(Level, const char* file, int line, std::string message)
{
    if (getenv("REF_SHIM_VERBOSE")) {
        fprintf(stderr, "[ref-gemm-shim] %s:%d %s\n", file ? file : "?", line, message.c_str());
    }
}
}  // namespace turbomind::core

using namespace turbomind;
using namespace turbomind::gemm;

namespace {

struct RefLinear {
    int          K{}, N{};
    void*        weight{};  // packed u4 (same byte count as the input)
    void*        scales{};  // packed (scale, -zero*scale) pairs
    MatrixLayout k_desc{};
    MatrixLayout q_desc{};
};

struct RefGemm {
    Gemm      gemm;
    Workspace ws{};
};

#define REF_CUDA(x)                                                                                                    \
    do {                                                                                                               \
        cudaError_t e_ = (x);                                                                                          \
        if (e_ != cudaSuccess) {                                                                                       \
            fprintf(stderr, "[ref-gemm-shim] %s: %s\n", #x, cudaGetErrorString(e_));                                   \
            return (int)e_;                                                                                            \
        }                                                                                                              \
    } while (0)

}  // namespace

extern "C" {

const char* ref_gemm_version()
{
    return "turbomind gemm::Gemm (sm80 s16816 kernels, compute_80 PTX -> sm_100 SASS) via oracle/ref_build/ref_gemm_shim.cu";
}

// One Gemm object (kernel registry, tuner, dispatch cache) + the workspace LlamaLinear::Impl::Impl allocates
// (LlamaLinear.cu:28-46).
int ref_gemm_create(void** out)
{
    auto g = std::make_unique<RefGemm>();
    g->ws.barriers_size   = Gemm::kBarriersSize;
    g->ws.partials_size   = Gemm::kPartialsSize;
    g->ws.tensormaps_size = 8192 * 128;
    REF_CUDA(cudaMalloc(&g->ws.barriers, g->ws.barriers_size));
    REF_CUDA(cudaMalloc(&g->ws.partials, g->ws.partials_size));
    REF_CUDA(cudaMalloc(&g->ws.tensormaps, g->ws.tensormaps_size));
    REF_CUDA(cudaMemset(g->ws.barriers, 0, g->ws.barriers_size));
    REF_CUDA(cudaMalloc((void**)&g->ws.flags, sizeof(int)));
    *out = g.release();
    return 0;
}

void ref_gemm_destroy(void* h)
{
    auto* g = (RefGemm*)h;
    if (g) {
        cudaFree(g->ws.barriers), cudaFree(g->ws.partials), cudaFree(g->ws.tensormaps), cudaFree(g->ws.flags);
        delete g;
    }
}

// qweight: int32 [K, N/8] (TurboMind pre-prepare u4 layout, nibble i of word w = column 8w+i), scales / zeros: half
// [K/group, N].  All device pointers; the inputs are not modified.
int ref_linear_create(void** out, const void* qweight, const void* scales, const void* zeros, int K, int N, int group,
                      void* stream)
{
    auto st = (cudaStream_t)stream;
    auto L  = std::make_unique<RefLinear>();
    L->K = K, L->N = N;
    const DataType data_type = kHalf;
    auto [conv_w, conv_s]    = GetConverters(data_type, kUint4, data_type, /*grouped=*/false, getSMVersion());
    if (!conv_w || !conv_s) {
        fprintf(stderr, "[ref-gemm-shim] no converter for u4 on sm %d\n", getSMVersion());
        return -1;
    }
    const size_t n_w = (size_t)K * N;
    // ---- weight (linear_weight.cc:196-253) ----
    uint16_t* tmp = nullptr;
    REF_CUDA(cudaMalloc((void**)&tmp, n_w * sizeof(uint16_t)));
    extend_to_u16(tmp, (const uint4_t*)qweight, n_w, st);
    if (conv_w->order == kRowMajor) {
        fprintf(stderr, "[ref-gemm-shim] row-major weight converter (pre-sm80 layout) is not handled\n");
        return -2;
    }
    MatrixLayout w_desc{data_type, conv_w->order, N, K, N};
    if (get_operand_tag(conv_w->pack) != OPERAND_A) {  // operand B: describe the same memory as [K, N]
        std::swap(w_desc.rows, w_desc.cols);
        w_desc.order = ~w_desc.order;
    }
    MatrixLayout kd = w_desc;
    kd.type         = data_type_v<uint4_t>;
    kd.pack         = conv_w->pack;
    REF_CUDA(cudaMalloc(&L->weight, n_w / 2));
    REF_CUDA(cudaMemsetAsync(L->weight, 0, n_w / 2, st));
    if (int ec = conv_w->Convert(tmp, w_desc, L->weight, kd, st)) {
        fprintf(stderr, "[ref-gemm-shim] weight Convert failed: %d\n", ec);
        return -3;
    }
    kd.type   = kUint4;
    L->k_desc = kd;
    // ---- scales + zeros (linear_weight.cc:255-312) ----
    const size_t n_s   = (size_t)(K / group) * N;
    half*        fused = nullptr;
    REF_CUDA(cudaMalloc((void**)&fused, n_s * 2 * sizeof(half)));
    fuse_scales_and_zeros(fused, (const half*)scales, (half*)zeros, n_s, st);
    MatrixLayout s_desc{kUint32, conv_s->order, N, K / group, N};
    if (get_operand_tag(conv_s->pack) != OPERAND_U) {  // operand V
        std::swap(s_desc.rows, s_desc.cols);
        s_desc.order = ~s_desc.order;
    }
    MatrixLayout qd = s_desc;
    qd.pack         = conv_s->pack;
    REF_CUDA(cudaMalloc(&L->scales, n_s * 2 * sizeof(half)));
    if (int ec = conv_s->Convert(fused, s_desc, L->scales, qd, st)) {
        fprintf(stderr, "[ref-gemm-shim] scale Convert failed: %d\n", ec);
        return -4;
    }
    L->q_desc = qd;
    REF_CUDA(cudaStreamSynchronize(st));
    REF_CUDA(cudaFree(tmp));
    REF_CUDA(cudaFree(fused));
    *out = L.release();
    return 0;
}

void ref_linear_destroy(void* h)
{
    auto* L = (RefLinear*)h;
    if (L) {
        cudaFree(L->weight), cudaFree(L->scales);
        delete L;
    }
}

// y[m, N or N/2] = x[m, K] . dequant(W); gated_silu != 0 runs the fused epilogue over interleaved (gate, up) columns.
// dispatch: 0 default heuristic, 1 measure (tune this shape on first use), 2 reuse.
int ref_linear_forward(void* gemm, void* linear, const void* x, int m, void* y, int gated_silu, int group, int dispatch,
                       void* stream)
{
    auto* g = (RefGemm*)gemm;
    auto* L = (RefLinear*)linear;
    Operation op{};
    op.dispatch  = (DispatchPolicy)dispatch;
    op.epilogue  = gated_silu ? Epilogue::kGatedSilu : Epilogue::kNone;
    op.quant_a   = QuantDesc{QuantType::kNone, 0};
    op.quant_b   = QuantDesc{QuantType::kK, group};
    op.batch_dim = 0;
    MatrixLayout desc_A{kHalf, kRowMajor, m, L->K, L->K};
    MatrixLayout desc_U{};
    const int    out_ld = gated_silu ? L->N / 2 : L->N;
    MatrixLayout desc_D{kHalf, kRowMajor, m, L->N, out_ld};
    MatrixLayout desc_W{};
    return g->gemm.Run(op, 1.f, x, desc_A, nullptr, desc_U, L->weight, L->k_desc, L->scales, L->q_desc, 0.f, y, desc_D, y,
                       desc_D, nullptr, desc_W, g->ws, (cudaStream_t)stream);
}

}  // extern "C"
