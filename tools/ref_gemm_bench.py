"""Development tool (measurement only): the reference's own W4A16 GEMM (oracle/_ref/libtm_ref_gemm.so -- gemm::Gemm::Run with
its sm80 s16816 kernels, tuned per shape by its own measurer) timed next to ours on the decode shapes, both captured in a
CUDA graph over rotating weight copies so every launch streams its weights from HBM.

    python tools/ref_gemm_bench.py [m=64] [iters=20]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lmdeploy_b200 import _lib as L            # noqa: E402
from oracle import ref_gemm_lib as RG         # noqa: E402

SHAPES = [("qkv", 4096, 6144, 0), ("wo", 4096, 4096, 0), ("w1w3", 4096, 28672, 1), ("w2", 14336, 4096, 0)]


def graph_time(fn, iters):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i + 3)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def measure(m=64, iters=20, shapes=SHAPES, verbose=True):
    """-> {name: dict(reference_us, ours_us, weight_bytes, max_abs_diff)} for the given (name, K, N, gated_silu) shapes."""
    lib = L.load()
    rg = RG.RefGemm()
    gen = torch.Generator(device="cuda").manual_seed(0)
    results = {}
    for name, K, N, epi in shapes:
        copies = max(2, (400 << 20) // (K * N // 2))
        x = torch.randn(m, K, generator=gen, device="cuda").half()
        qs, ss, zs, ref_lin, ours = [], [], [], [], []
        for c in range(copies):
            qw = torch.randint(-2 ** 31, 2 ** 31 - 1, (K, N // 8), generator=gen, device="cuda", dtype=torch.int32)
            sc = ((torch.rand(K // 128, N, generator=gen, device="cuda") + 0.5) * (0.02 / K ** 0.5)).half()
            zp = torch.randint(6, 11, (K // 128, N), generator=gen, device="cuda").half()
            ref_lin.append(RG.RefLinear(rg, qw, sc, zp, K, N, torch))
            d = L.LinearDesc(K, N, 1, 0, 128, None, None)
            prepared = torch.empty(lib.b200_linear_prepared_bytes(C.byref(d)), dtype=torch.uint8, device="cuda")
            d.prepared = prepared.data_ptr()
            L.check(lib.b200_linear_prepare(C.byref(d), qw.data_ptr(), sc.data_ptr(), zp.data_ptr(), 0, prepared.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "prepare")
            ours.append((d, prepared))
            del qw, sc, zp
        n_out = N // 2 if epi else N
        y_ref = torch.empty(m, n_out, dtype=torch.float16, device="cuda")
        y_our = torch.empty(m, n_out, dtype=torch.float16, device="cuda")
        ws = torch.zeros(lib.b200_linear_workspace_bytes(C.byref(ours[0][0]), 128), dtype=torch.uint8, device="cuda")
        # let the reference tune this shape with its own measurer, then replay its choice
        ref_lin[0].forward(x, gated_silu=bool(epi), dispatch=RG.DISPATCH_MEASURE, out=y_ref)
        torch.cuda.synchronize()

        def run_ref(i):
            ref_lin[i % copies].forward(x, gated_silu=bool(epi), dispatch=RG.DISPATCH_REUSE, out=y_ref)

        # ours runs as it does inside the decode step: activations already in the packed K-major layout
        mt = (m + 15) // 16 * 16
        xp = torch.zeros(K // 8, mt, 8, dtype=torch.float16, device="cuda")
        L.check(lib.b200_pack_activations(x.data_ptr(), K, m, K, mt, 0, xp.data_ptr(), torch.cuda.current_stream().cuda_stream),
                "pack")

        def run_ours(i):
            d = ours[i % copies][0]
            L.check(lib.b200_linear_forward_packed(C.byref(d), xp.data_ptr(), mt, m, y_our.data_ptr(), n_out, epi, 0,
                                                   ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream),
                    "forward_packed")

        t_ref, t_our = graph_time(run_ref, iters), graph_time(run_ours, iters)
        run_ref(0), run_ours(0)
        torch.cuda.synchronize()
        # (different random weights per copy; copy 0 on both sides)
        diff = (y_ref.float() - y_our.float()).abs().max().item()
        wb = K * N / 2 + K // 128 * N * 4
        results[name] = dict(K=K, N=N, reference_us=t_ref, ours_us=t_our, weight_bytes=int(wb), max_abs_diff=diff)
        if verbose:
            print(f"{name:5s} K={K:5d} N={N:5d} m={m}: reference {t_ref:7.2f} us ({wb / t_ref * 1e-3:6.0f} GB/s)   "
                  f"ours {t_our:7.2f} us ({wb / t_our * 1e-3:6.0f} GB/s)   ours/ref speed-up {t_ref / t_our:4.2f}x   "
                  f"max|y_ref - y_ours| {diff:.3g}")
        for r in ref_lin:
            r.close()
    rg.close()
    return results


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    if not RG.available():
        print("oracle/_ref/libtm_ref_gemm.so is not built")
        return 1
    print(RG.load().ref_gemm_version().decode())
    measure(m, iters)
    return 0


if __name__ == "__main__":
    sys.exit(main())
