// Tensor-parallel communicator over NVLink peer memory: fused all-reduce + residual + RMSNorm in ONE kernel.
//
// Replaces DeviceCommImpl::AllreduceResidualBiasRMSnorm (src/turbomind/comm/device_comm.h:88-101; native
// kernels comm/cuda_ipc/fused_allreduce.cu:25-553) for a one-process-per-GPU world.  Design (B200 / NVSwitch):
// messages are 0.5-0.8 MB, latency-bound, and NVSwitch gives every peer full bandwidth, so each rank
//   1. copies nothing: the producing GEMM wrote its partial [tokens, dim] tile into this rank's exchange buffer;
//   2. publishes "my partial for epoch e is ready" by writing a flag into every peer's flag array
//      (st.release.sys over NVLink);
//   3. waits until all peers' flags for epoch e arrived (ld.acquire.sys on its own flag array);
//   4. each CTA owns a slice of tokens: it pulls the same slice from every peer with 16-byte peer loads, sums
//      the ranks in rank order in T (bit-identical on every rank), adds the residual, reduces sum-of-squares,
//      normalises and writes the next GEMM's input tile;
//   5. no second ("done reading") flag round: the two exchanges of a layer use two different halves of the exchange
//      buffer (b200_comm_local_buffer(comm, 0 / 1)), so a half is rewritten by a GEMM only after the NEXT exchange's
//      arrival round, and arriving there implies having finished the reads of the previous one.
// Flags are monotonically increasing epochs, so no reset traffic is needed.  The epoch lives in DEVICE memory, one
// counter per CTA slot (every rank runs the same kernel sequence with the same grids, so slot i counts the same
// on all ranks): a captured CUDA graph of the step replays correctly.
#include "../../include/b200_turbomind.h"
#include "common.cuh"
#include "host_utils.h"

namespace b200 {
namespace comm {

constexpr int kMaxRanks  = 8;
constexpr int kFlagSlots = 128;  // one slot per CTA; 512-byte line per (phase, src-rank)
constexpr int kBlock     = 256;

struct Peers {
    void* data[kMaxRanks];
    int*  flags[kMaxRanks];
    int*  slot_epoch;  // [kFlagSlots] this rank's per-CTA-slot epoch counters (device memory)
    void* mc_data;     // NVLS: multicast address of the (symmetric) exchange buffers, or null
};

// epoch of this launch for this CTA slot (read by thread 0, broadcast through shared memory)
__device__ __forceinline__ int next_epoch(const Peers& peers)
{
    __shared__ int s_epoch;
    if (threadIdx.x == 0) {
        s_epoch = (int)((unsigned)peers.slot_epoch[blockIdx.x] + 1u);  // wraps modulo 2^32 (epoch_before)
    }
    __syncthreads();
    return s_epoch;
}
__device__ __forceinline__ void commit_epoch(const Peers& peers, int epoch)
{
    if (threadIdx.x == 0) {
        peers.slot_epoch[blockIdx.x] = epoch;
    }
}

__device__ __forceinline__ void st_release_sys(int* p, int v)
{
    asm volatile("st.release.sys.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p)
{
    int v;
    asm volatile("ld.acquire.sys.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Flag traffic of the arrival round.  The payload a flag announces was written by the PREVIOUS kernel of the stream
// (complete and in L2 before this kernel started), so the store needs no release fence (a system-scope release makes
// the SM drain every outstanding access first); the payload reads that follow are ld.volatile (never L1-cached) and
// are ordered after the spin by bar.sync.
__device__ __forceinline__ void st_relaxed_sys(int* p, int v)
{
    asm volatile("st.relaxed.sys.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_sys(const int* p)
{
    int v;
    asm volatile("ld.relaxed.sys.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// 16-byte load that never hits a stale L1 line (peer memory is not coherent with the local L1)
__device__ __forceinline__ uint4 ld_peer_v4(const void* p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// NVLS: one load returns the sum over every rank's copy of the address, reduced inside the NVSwitch
// (multimem.ld_reduce, the instruction the reference's NVLS path uses: comm/cuda_ipc/multimem.cuh:14-45).
template<class T>
__device__ __forceinline__ uint4 multimem_ld_reduce_v4(const void* mc_ptr)
{
    uint4 v;
    if constexpr (std::is_same<T, half>::value) {
        asm volatile("multimem.ld_reduce.weak.global.add.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(mc_ptr)
                     : "memory");
    }
    else {
        asm volatile("multimem.ld_reduce.weak.global.add.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(mc_ptr)
                     : "memory");
    }
    return v;
}

// Epochs grow by one per launch for the life of the communicator (2 L + 1 launches per step): compare them modulo
// 2^32 so the waits stay correct when the counter wraps (a plain `flag < epoch` falls through at INT_MIN).
__device__ __forceinline__ bool epoch_before(int flag, int epoch)
{
    return (int)((unsigned)flag - (unsigned)epoch) < 0;
}

// flag layout in every rank's flag array: [phase(2)][src rank(8)][kFlagSlots]
__device__ __forceinline__ int* flag_ptr(int* base, int phase, int src)
{
    return base + ((size_t)phase * kMaxRanks + src) * kFlagSlots;
}

// signal all peers + wait for all peers (called by one thread per CTA 0 .. then broadcast via grid?)
// Every CTA performs its own signalling on its own flag slot (slot = blockIdx.x) so no grid sync is needed.
// arrival round for payloads produced by an earlier kernel (see st_relaxed_sys)
__device__ __forceinline__ void arrive_phase_relaxed(const Peers& peers, int rank, int world, int epoch)
{
    if (threadIdx.x < world) {
        const int r = threadIdx.x;
        st_relaxed_sys(flag_ptr(peers.flags[r], 0, rank) + blockIdx.x, epoch);
        const int* f = flag_ptr(peers.flags[rank], 0, r) + blockIdx.x;
        while (epoch_before(ld_relaxed_sys(f), epoch)) {}
        // the relaxed spin + bar.sync give no system-scope acquire under the PTX memory model: one fence by the
        // spinning threads orders the payload loads of the whole CTA (after the barrier) behind the observed flags
        asm volatile("fence.acquire.sys;" ::: "memory");
    }
    __syncthreads();
}

__device__ __forceinline__ void barrier_phase(const Peers& peers, int rank, int world, int phase, int epoch)
{
    if (threadIdx.x < world) {
        const int r = threadIdx.x;
        st_release_sys(flag_ptr(peers.flags[r], phase, rank) + blockIdx.x, epoch);
    }
    if (threadIdx.x < world) {
        const int* f = flag_ptr(peers.flags[rank], phase, threadIdx.x) + blockIdx.x;
        while (epoch_before(ld_acquire_sys(f), epoch)) {}
    }
    __syncthreads();
}

constexpr int kArBlock   = 512;
constexpr int kArMaxPass = 2;  // dim <= 512 * 8 * 2 = 8192

// One CTA per token (grid-strided).  All peer loads of a token (world x passes 16-byte loads per thread) are issued
// before the first add, so the exchange costs ONE NVLink round trip; the sum runs in rank order in T (bit-identical
// on every rank), then residual add, sum of squares, normalisation -- the row never leaves registers.
// MC = true: the rank sum comes from ONE multimem.ld_reduce per 16 bytes (in-switch reduction, NVLS) instead of `world` peer
// loads: a rank then moves tokens * dim * 2 bytes over its links instead of world times that.
template<class T, bool MC>
__global__ void __launch_bounds__(kArBlock) allreduce_residual_rmsnorm_kernel(
    Peers peers, int rank, int world, T* out, int out_mt, T* residual, const T* weight, const T* bias,
    int tokens, int dim, float eps, size_t data_offset)
{
    pdl_launch_dependents();
    pdl_wait();
    using Ops = TypeOps<T>;
    const int epoch = next_epoch(peers);
    // the partial tile was written by the previous kernel on this stream (complete => in L2, where peers read it)
    arrive_phase_relaxed(peers, rank, world, epoch);

    __shared__ float s_part[kArBlock / 32];
    __shared__ float s_total;
    const int npass = (dim + kArBlock * 8 - 1) / (kArBlock * 8);
    for (int tok = blockIdx.x; tok < tokens; tok += gridDim.x) {
        T*    r = residual + (size_t)tok * dim;
        uint4 v[kArMaxPass][kMaxRanks];
#pragma unroll
        for (int ps = 0; ps < kArMaxPass; ++ps) {
            const int i = (ps * kArBlock + threadIdx.x) * 8;
            if constexpr (MC) {
                if (ps < npass && i < dim) {
                    v[ps][0] = multimem_ld_reduce_v4<T>(reinterpret_cast<const T*>((const char*)peers.mc_data + data_offset)
                                                        + (size_t)tok * dim + i);
                }
            }
            else {
#pragma unroll
                for (int q = 0; q < kMaxRanks; ++q) {
                    if (ps < npass && q < world && i < dim) {
                        v[ps][q] = ld_peer_v4(reinterpret_cast<const T*>((const char*)peers.data[q] + data_offset)
                                              + (size_t)tok * dim + i);
                    }
                }
            }
        }
        T     rv[kArMaxPass][8];
        float acc = 0.f;
#pragma unroll
        for (int ps = 0; ps < kArMaxPass; ++ps) {
            const int i = (ps * kArBlock + threadIdx.x) * 8;
            if (ps < npass && i < dim) {
                T sum[8];
                *reinterpret_cast<uint4*>(&sum[0]) = v[ps][0];
#pragma unroll
                for (int q = 1; q < kMaxRanks; ++q) {
                    if (!MC && q < world) {
                        T pv[8];
                        *reinterpret_cast<uint4*>(&pv[0]) = v[ps][q];
#pragma unroll
                        for (int j = 0; j < 8; ++j) sum[j] = Ops::add(sum[j], pv[j]);
                    }
                }
                *reinterpret_cast<uint4*>(&rv[ps][0]) = *reinterpret_cast<const uint4*>(r + i);
#pragma unroll
                for (int j = 0; j < 8; ++j) rv[ps][j] = Ops::add(rv[ps][j], sum[j]);
                if (bias) {
                    T bv[8];
                    *reinterpret_cast<uint4*>(&bv[0]) = *reinterpret_cast<const uint4*>(bias + i);
#pragma unroll
                    for (int j = 0; j < 8; ++j) rv[ps][j] = Ops::add(rv[ps][j], bv[j]);
                }
                *reinterpret_cast<uint4*>(r + i) = *reinterpret_cast<const uint4*>(&rv[ps][0]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float f = Ops::to_f32(rv[ps][j]);
                    acc           = fmaf(f, f, acc);
                }
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < kArBlock / 32; ++i) t += s_part[i];
            s_total = t;
        }
        __syncthreads();
        const float inv_rms = rsqrtf(s_total / (float)dim + eps);
#pragma unroll
        for (int ps = 0; ps < kArMaxPass; ++ps) {
            const int i = (ps * kArBlock + threadIdx.x) * 8;
            if (ps < npass && i < dim) {
                T wv[8], ov[8];
                *reinterpret_cast<uint4*>(&wv[0]) = *reinterpret_cast<const uint4*>(weight + i);
#pragma unroll
                for (int j = 0; j < 8; ++j) ov[j] = Ops::mul(Ops::from_f32(Ops::to_f32(rv[ps][j]) * inv_rms), wv[j]);
                T* dst = out_mt > 0 ? out + ((size_t)(i >> 3) * out_mt + tok) * 8 : out + (size_t)tok * dim + i;
                *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(&ov[0]);
            }
        }
        __syncthreads();
    }
    commit_epoch(peers, epoch);
}

template<class T>
__global__ void __launch_bounds__(kBlock) allreduce_sum_kernel(Peers peers, int rank, int world, T* data,
                                                               size_t count, size_t data_offset)
{
    pdl_launch_dependents();
    pdl_wait();
    using Ops = TypeOps<T>;
    const int epoch = next_epoch(peers);
    // stage: copy my data into my exchange buffer
    T* mine = reinterpret_cast<T*>((char*)peers.data[rank] + data_offset);
    for (size_t i = (blockIdx.x * (size_t)kBlock + threadIdx.x) * 8; i < count; i += (size_t)gridDim.x * kBlock * 8) {
        *reinterpret_cast<uint4*>(mine + i) = *reinterpret_cast<const uint4*>(data + i);
    }
    __threadfence_system();
    __syncthreads();
    barrier_phase(peers, rank, world, 0, epoch);
    for (size_t i = (blockIdx.x * (size_t)kBlock + threadIdx.x) * 8; i < count; i += (size_t)gridDim.x * kBlock * 8) {
        T sum[8];
        for (int q = 0; q < world; ++q) {
            T v[8];
            *reinterpret_cast<uint4*>(&v[0]) =
                ld_peer_v4(reinterpret_cast<const T*>((const char*)peers.data[q] + data_offset) + i);
#pragma unroll
            for (int j = 0; j < 8; ++j) sum[j] = q == 0 ? v[j] : Ops::add(sum[j], v[j]);
        }
        *reinterpret_cast<uint4*>(data + i) = *reinterpret_cast<const uint4*>(&sum[0]);
    }
    barrier_phase(peers, rank, world, 1, epoch);
    commit_epoch(peers, epoch);
}

// vocab-sharded greedy: local argmax (one CTA per row), exchange (value, index), pick global winner
template<class T>
__global__ void __launch_bounds__(1024) argmax_exchange_kernel(Peers peers, int rank, int world, int* out,
                                                               const T* logits, int rows, int vocab, int valid,
                                                               size_t data_offset)
{
    pdl_launch_dependents();
    pdl_wait();
    using Ops = TypeOps<T>;
    const int epoch = next_epoch(peers);
    __shared__ float s_v[32];
    __shared__ int   s_i[32];
    const int row = blockIdx.x;
    const T*  lp  = logits + (size_t)row * vocab;
    float     bv  = -INFINITY;
    int       bi  = 0x7fffffff;
    // only the first `valid` columns of this rank's shard are real vocabulary entries; the pad columns up to `vocab`
    // hold logit 0 (zero lm_head rows) and must not win when every real logit is negative (sampling.cc:132 passes
    // vocab_size next to vocab_size_padded for the same reason)
    for (int i = threadIdx.x; i < valid; i += blockDim.x) {
        const float v = Ops::to_f32(lp[i]);
        if (v > bv || (v == bv && i < bi)) {
            bv = v, bi = i;
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, m);
        const int   oi = __shfl_xor_sync(0xffffffffu, bi, m);
        if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov, bi = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        s_v[threadIdx.x >> 5] = bv, s_i[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) {
            if (s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi)) {
                bv = s_v[w], bi = s_i[w];
            }
        }
        // publish into my exchange buffer: [rows][2] (value bits, global index)
        int* mine = reinterpret_cast<int*>((char*)peers.data[rank] + data_offset);
        mine[2 * row]     = __float_as_int(bv);
        mine[2 * row + 1] = rank * vocab + bi;
    }
    __threadfence_system();
    __syncthreads();
    // per-row CTAs: slot = blockIdx.x (< kFlagSlots rows supported per launch chunk)
    barrier_phase(peers, rank, world, 0, epoch);
    if (threadIdx.x == 0) {
        float gv = -INFINITY;
        int   gi = 0x7fffffff;
        for (int q = 0; q < world; ++q) {
            const int*  src = reinterpret_cast<const int*>((const char*)peers.data[q] + data_offset);
            const float v   = __int_as_float(src[2 * row]);
            const int   idx = src[2 * row + 1];
            if (v > gv || (v == gv && idx < gi)) {
                gv = v, gi = idx;
            }
        }
        out[row] = gi;
    }
    barrier_phase(peers, rank, world, 1, epoch);
    commit_epoch(peers, epoch);
}

}  // namespace comm
}  // namespace b200

using namespace b200;
using namespace b200::comm;

struct b200_comm {
    int    rank{}, world{};
    Peers  peers{};
    size_t buf_bytes{};
    size_t half_bytes{};  // offset of the second half of the exchange buffer
};

extern "C" size_t b200_comm_buffer_bytes(int max_tokens, int hidden, int vocab_per_rank)
{
    (void)vocab_per_rank;
    size_t a = ((size_t)max_tokens * hidden * 2 + 4095) / 4096 * 4096;
    return 2 * a + 4096;  // two halves (see b200_comm_local_buffer) + the argmax exchange page
}

extern "C" size_t b200_comm_flags_bytes(void)
{
    return (size_t)2 * kMaxRanks * kFlagSlots * sizeof(int);
}

extern "C" int b200_comm_set_geometry(b200_comm* c, int max_tokens, int hidden)
{
    if (!c || max_tokens <= 0 || hidden <= 0) {
        return fail(B200_ERR_INVALID_ARG, "comm_set_geometry: bad arguments");
    }
    c->half_bytes = ((size_t)max_tokens * hidden * 2 + 4095) / 4096 * 4096;
    return 0;
}

extern "C" int b200_comm_create(b200_comm** out, int rank, int world, void* const* peer_data, void* const* peer_flags)
{
    if (!out || !peer_data || !peer_flags || world < 1 || world > kMaxRanks || rank < 0 || rank >= world) {
        return fail(B200_ERR_INVALID_ARG, "comm_create: bad arguments");
    }
    auto* c  = new b200_comm();
    c->rank  = rank;
    c->world = world;
    for (int r = 0; r < world; ++r) {
        c->peers.data[r]  = peer_data[r];
        c->peers.flags[r] = (int*)peer_flags[r];
    }
    cudaError_t e = cudaMalloc((void**)&c->peers.slot_epoch, kFlagSlots * sizeof(int));
    if (e == cudaSuccess) {
        e = cudaMemset(c->peers.slot_epoch, 0, kFlagSlots * sizeof(int));
    }
    if (e != cudaSuccess) {
        delete c;
        return fail_cuda(e, "comm_create: epoch counters");
    }
    *out = c;
    return 0;
}

extern "C" int b200_comm_set_multicast(b200_comm* c, void* mc_data)
{
    if (!c) {
        return fail(B200_ERR_INVALID_ARG, "comm_set_multicast: comm is NULL");
    }
    c->peers.mc_data = mc_data;  // null switches back to the peer-load kernel
    return 0;
}

extern "C" void* b200_comm_local_buffer(b200_comm* c, int half)
{
    return c ? (char*)c->peers.data[c->rank] + (half ? c->half_bytes : 0) : nullptr;
}

extern "C" void b200_comm_destroy(b200_comm* c)
{
    if (c && c->peers.slot_epoch) {
        cudaFree(c->peers.slot_epoch);
    }
    delete c;
}

static int comm_grid(int tokens)
{
    return std::max(1, std::min(tokens, kFlagSlots));
}

extern "C" int b200_comm_allreduce_residual_rmsnorm(b200_comm* c, void* out, int out_packed_mt, void* residual,
                                                    const void* hidden_partial, const void* weight, const void* bias,
                                                    int tokens, int dim, float eps, int dtype, void* stream)
{
    if (!c || !out || !residual || !hidden_partial || !weight || tokens <= 0 || dim % 8 || dim > kArBlock * 8 * kArMaxPass) {
        return fail(B200_ERR_INVALID_ARG, "allreduce_residual_rmsnorm: bad arguments");
    }
    // the partial must live inside this rank's exchange buffer
    const char* base = (const char*)c->peers.data[c->rank];
    const size_t off = (const char*)hidden_partial - base;
    cudaStream_t st  = (cudaStream_t)stream;
    const int    grid = comm_grid(tokens);
    const bool mc = c->peers.mc_data != nullptr;
    if (dtype == B200_DTYPE_F16) {
        auto kern = mc ? allreduce_residual_rmsnorm_kernel<half, true> : allreduce_residual_rmsnorm_kernel<half, false>;
        launch_k(kern, dim3(grid), dim3(kArBlock), 0, st, c->peers, c->rank, c->world, (half*)out, out_packed_mt, (half*)residual,
                 (const half*)weight, (const half*)bias, tokens, dim, eps, off);
    }
    else {
        auto kern = mc ? allreduce_residual_rmsnorm_kernel<nv_bfloat16, true> : allreduce_residual_rmsnorm_kernel<nv_bfloat16, false>;
        launch_k(kern, dim3(grid), dim3(kArBlock), 0, st, c->peers, c->rank, c->world, (nv_bfloat16*)out, out_packed_mt,
                 (nv_bfloat16*)residual, (const nv_bfloat16*)weight, (const nv_bfloat16*)bias, tokens, dim, eps, off);
    }
    return check_launch("allreduce_residual_rmsnorm_kernel");
}

extern "C" int b200_comm_allreduce_sum(b200_comm* c, void* data, size_t count, int dtype, void* stream)
{
    if (!c || !data || count % 8) {
        return fail(B200_ERR_INVALID_ARG, "allreduce_sum: bad arguments");
    }
    const int grid = (int)std::min<size_t>(kFlagSlots, (count / 8 + kBlock - 1) / kBlock);
    if (dtype == B200_DTYPE_F16) {
        launch_k(allreduce_sum_kernel<half>, dim3(grid), dim3(kBlock), 0, (cudaStream_t)stream, c->peers, c->rank, c->world,
                                                                              (half*)data, count, 0);
    }
    else {
        launch_k(allreduce_sum_kernel<nv_bfloat16>, dim3(grid), dim3(kBlock), 0, (cudaStream_t)stream, c->peers, c->rank, c->world,
                                                                                     (nv_bfloat16*)data, count, 0);
    }
    return check_launch("allreduce_sum_kernel");
}

extern "C" int b200_comm_argmax_allgather(b200_comm* c, int* out_tokens, const void* local_logits, int rows,
                                          int vocab_per_rank, int valid_per_rank, int dtype, void* stream)
{
    if (!c || !out_tokens || !local_logits || rows <= 0 || rows > kFlagSlots || valid_per_rank < 0
        || valid_per_rank > vocab_per_rank) {
        return fail(B200_ERR_INVALID_ARG, "argmax_allgather: bad arguments (rows <= 128)");
    }
    // (value, index) pairs live in the page behind the two exchange halves
    const size_t off = 2 * c->half_bytes;
    if (dtype == B200_DTYPE_F16) {
        launch_k(argmax_exchange_kernel<half>, dim3(rows), dim3(1024), 0, (cudaStream_t)stream, c->peers, c->rank, c->world, out_tokens,
                                                                              (const half*)local_logits, rows,
                                                                              vocab_per_rank, valid_per_rank, off);
    }
    else {
        launch_k(argmax_exchange_kernel<nv_bfloat16>, dim3(rows), dim3(1024), 0, (cudaStream_t)stream, 
            c->peers, c->rank, c->world, out_tokens, (const nv_bfloat16*)local_logits, rows, vocab_per_rank,
            valid_per_rank, off);
    }
    return check_launch("argmax_exchange_kernel");
}

// ---- step advance (device-side bookkeeping so a captured graph can iterate) ----
__global__ void step_advance_kernel(int* token_ids, const int* out_tokens, int* cu_k_len, const uint8_t* finished, int batch)
{
    pdl_launch_dependents();
    pdl_wait();
    // single block: k_len[b] += 1 for unfinished sequences -> rebuild prefix
    __shared__ int s_len[1024];
    for (int b = threadIdx.x; b < batch; b += blockDim.x) {
        const bool fin = finished && finished[b];
        s_len[b]       = cu_k_len[b + 1] - cu_k_len[b] + (fin ? 0 : 1);
        if (!fin) {
            token_ids[b] = out_tokens[b];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < batch; ++b) {
            acc += s_len[b];
            cu_k_len[b + 1] = acc;
        }
    }
}

extern "C" int b200_step_advance(const int* token_ids, const int* out_tokens, int* cu_k_len, const uint8_t* finished,
                                 int batch, void* stream)
{
    if (!token_ids || !out_tokens || !cu_k_len || batch <= 0 || batch > 1024) {
        return fail(B200_ERR_INVALID_ARG, "step_advance: bad arguments");
    }
    launch_k(step_advance_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, (int*)token_ids, out_tokens, cu_k_len, finished, batch);
    return check_launch("step_advance_kernel");
}
