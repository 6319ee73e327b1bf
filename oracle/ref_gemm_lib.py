"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_ref/libtm_ref_gemm.so: the REFERENCE's own weight-only-quant GEMM
(src/turbomind/kernels/gemm: gemm::Gemm::Run + the sm80 s16816 kernel family, weight repack through the reference's
converters) compiled in place by oracle/ref_build/Makefile (kernel bodies from compute_80 PTX -> sm_100 SASS, see
ref_gemm_shim.cu).

Only tests/ and explicitly requested measurement legs may import this module; the product path (lmdeploy_b200/) never
does.  The .so is built where /root/reference exists and travels to the GPU box with the repo snapshot."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.abspath(__file__))
# B200_REF_GEMM_LIB=ptx: the build whose kernel TUs hold compute_80 PTX only (JIT-compiled by the driver at load time, the way
# the reference's wheels reach Blackwell); default: the ahead-of-time sm_100 SASS build of the same sources
# default ("dd"): the reference's kernel templates registered with the evict-first L2 hint off (oracle/ref_build/
# ref_gemm_kernels_dd.cu) -- the reference's own registration raises "illegal instruction" on a B200 with this toolchain;
# "sass" = its own registration compiled ahead of time, "ptx" = the same through the driver JIT (both kept for the record)
LIB_PATH = os.path.join(ROOT, "_ref", {"ptx": "libtm_ref_gemm_ptx.so", "sass": "libtm_ref_gemm.so"}.get(
    os.environ.get("B200_REF_GEMM_LIB", "dd"), "libtm_ref_gemm_dd.so"))
DISPATCH_DEFAULT, DISPATCH_MEASURE, DISPATCH_REUSE = 0, 1, 2

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_gemm_version.restype = C.c_char_p
        _lib.ref_gemm_create.argtypes = [C.POINTER(C.c_void_p)]
        _lib.ref_gemm_destroy.argtypes = [C.c_void_p]
        _lib.ref_linear_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.c_void_p]
        _lib.ref_linear_destroy.argtypes = [C.c_void_p]
        _lib.ref_linear_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                            C.c_int, C.c_void_p]
    return _lib


class RefGemm:
    """One reference Gemm object (registry + tuner + workspace)."""

    def __init__(self):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.ref_gemm_create(C.byref(self.h))
        if rc:
            raise RuntimeError(f"ref_gemm_create failed: {rc}")

    def close(self):
        if self.h:
            self.lib.ref_gemm_destroy(self.h)
            self.h = C.c_void_p()


class RefLinear:
    """A u4 (g128) linear prepared by the reference's converters from TurboMind pre-prepare tensors (torch CUDA tensors:
    qweight int32 [K, N/8], scales / zeros fp16 [K/128, N])."""

    def __init__(self, gemm: RefGemm, qweight, scales, zeros, K, N, torch, group=128):
        self.gemm, self.K, self.N, self.group, self.torch = gemm, K, N, group, torch
        self.h = C.c_void_p()
        zeros = zeros.clone()            # fuse_scales_and_zeros takes a non-const pointer
        st = torch.cuda.current_stream().cuda_stream
        rc = gemm.lib.ref_linear_create(C.byref(self.h), qweight.data_ptr(), scales.data_ptr(), zeros.data_ptr(), K, N,
                                        group, st)
        if rc:
            raise RuntimeError(f"ref_linear_create failed: {rc}")

    def forward(self, x, gated_silu=False, dispatch=DISPATCH_DEFAULT, out=None):
        torch = self.torch
        m = x.shape[0]
        n_out = self.N // 2 if gated_silu else self.N
        y = out if out is not None else torch.empty(m, n_out, dtype=torch.float16, device=x.device)
        rc = self.gemm.lib.ref_linear_forward(self.gemm.h, self.h, x.data_ptr(), m, y.data_ptr(), int(gated_silu),
                                              self.group, dispatch, torch.cuda.current_stream().cuda_stream)
        if rc:
            raise RuntimeError(f"reference Gemm::Run returned {rc}")
        return y

    def close(self):
        if self.h:
            self.gemm.lib.ref_linear_destroy(self.h)
            self.h = C.c_void_p()
