// TEST INFRASTRUCTURE -- kernel registration for oracle/_ref/libtm_ref_gemm_dd.so.
//
// The reference registers its W4A16 kernels for Blackwell in src/turbomind/kernels/gemm/kernel/sm80_16816_4.cu:16-57:
// instances of Config_U4_d<Sm80, half, kColMajor>::Type<M, N, K, wm, wn, wk, PolicyA, PolicyB, stages, ...>
// (arch/config_sm80_s16816.h:108-121).  Most of them stream the weights with PolicyB = cache_policy::Stream, i.e.
// cp.async ... .L2::cache_hint with a createpolicy.fractional.L2::evict_first descriptor (iterator_sm80.h:145-149,
// cp_async.h:127-137).  With this image's toolchain (ptxas 12.9.86, and the 580 driver's JIT on compute_80 PTX alike)
// that instruction is encoded as `LDGSTS... desc[UR1]` and a B200 raises "illegal instruction" on it
// (compute-sanitizer, profiles/r02_ref_gemm_illegal_instruction.md) -- the reference's own default dispatch picks such
// a kernel at m = 1 and m = 64, so its GEMM cannot run here as built.
//
// This translation unit registers the SAME templates at the SAME tile shapes with the evict-first hint switched off
// (PolicyB = cache_policy::Default): identical arithmetic (dequant, mma.sync order, split-K, epilogue), only the L2
// eviction hint of the weight stream differs.  It is compiled from compute_80 PTX to sm_100 SASS like the reference's
// own kernel TUs and linked INSTEAD of sm80_16816_4.cu; nothing of the reference is copied, its headers are used in place.
#include "src/turbomind/kernels/gemm/arch.h"
#include "src/turbomind/kernels/gemm/arch/config_sm80_s16816.h"
#include "src/turbomind/kernels/gemm/registrar.h"
#include "src/turbomind/kernels/gemm/types.h"
#include <type_traits>

namespace turbomind::gemm {

using namespace sm80_s16816;
using D = cache_policy::Default;

namespace {
Registrar reg_dd([](Collector& c, int /*arch*/) {
    constexpr int G = 128;  // AWQ group size on this path
    using C         = Config_U4_d<Sm80, half, kColMajor>;
    // tile shapes / warp maps / stage counts of sm80_16816_4.cu:23-56, PolicyB -> Default
    c.add<C::Type<128, 256, 32, 1, 8, 1, D, D, 3, true, 1, G, 128, 128>>();
    c.add<C::Type<128, 256, 32, 1, 8, 1, D, D, 4, true, 1, G, 128, 128>>();
    c.add<C::Type<128, 128, 32, 1, 4, 1, D, D, 3, true, 1, G, 64, 128>>();
    c.add<C::Type<128, 128, 32, 1, 4, 1, D, D, 4, true, 1, G, 64, 128>>();
    c.add<C::Type<128, 128, 64, 1, 4, 2, D, D, 3, true, 1, G, 64, 128>>();

    c.add<C::Type<96, 256, 32, 1, 8, 1, D, D, 4, true, 1, G>>();
    c.add<C::Type<96, 256, 32, 1, 8, 1, D, D, 3, true, 1, G>>();
    c.add<C::Type<96, 128, 32, 1, 4, 1, D, D, 4, true, 1, G>>();
    c.add<C::Type<96, 128, 128, 1, 4, 2, D, D, 3, true, 1, G>>();

    c.add<C::Type<64, 256, 32, 1, 4, 1, D, D, 3, true, 1, G, 64, 128>>();
    c.add<C::Type<64, 256, 32, 1, 4, 1, D, D, 4, true, 1, G, 64, 128>>();
    c.add<C::Type<64, 128, 32, 1, 4, 1, D, D, 4, true, 1, G>>();
    c.add<C::Type<64, 128, 64, 1, 4, 1, D, D, 3, true, 1, G>>();
    c.add<C::Type<64, 128, 128, 1, 4, 2, D, D, 3, true, 1, G>>();
    c.add<C::Type<64, 64, 64, 1, 2, 2, D, D, 6, true, 1, G>>();

    c.add<C::Type<48, 256, 64, 1, 4, 1, D, D, 3, true, 1, G, 48, 128>>();
    c.add<C::Type<48, 128, 64, 1, 4, 1, D, D, 4, true, 1, G>>();
    c.add<C::Type<48, 128, 128, 1, 4, 2, D, D, 3, true, 1, G>>();
    c.add<C::Type<48, 64, 128, 1, 2, 2, D, D, 4, true, 1, G>>();

    c.add<C::Type<32, 256, 64, 1, 4, 1, D, D, 3, true, 1, G>>();
    c.add<C::Type<32, 128, 64, 1, 4, 1, D, D, 4, true, 1, G>>();
    c.add<C::Type<32, 128, 128, 1, 4, 2, D, D, 3, true, 1, G>>();
    c.add<C::Type<32, 64, 128, 1, 2, 2, D, D, 3, true, 1, G>>();
    c.add<C::Type<32, 64, 128, 1, 2, 2, D, D, 4, true, 1, G>>();

    c.add<C::Type<16, 128, 64, 1, 4, 1, D, D, 4, true, 1, G>>();
    c.add<C::Type<16, 128, 128, 1, 4, 2, D, D, 3, true, 1, G>>();
    c.add<C::Type<16, 128, 128, 1, 4, 2, D, D, 4, true, 1, G>>();
    c.add<C::Type<16, 64, 128, 1, 2, 2, D, D, 3, true, 1, G>>();
    c.add<C::Type<16, 64, 128, 1, 2, 2, D, D, 4, true, 1, G>>();
});
}  // namespace

}  // namespace turbomind::gemm
