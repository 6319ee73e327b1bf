"""`_turbomind` -- the module name the reference's Python imports (lmdeploy/turbomind/turbomind.py:29-32,
tests/turbomind/linear/linear.py:11-16) -- implemented over this repository's C ABI (libb200tm.so).

Scope: the LINEAR surface of src/turbomind/python/linear_bind.cpp:95-294 and the tensor helpers of bind.cpp:661-740 that
tests/turbomind/linear/{linear,fixture}.py touch for DENSE weights (bf16 / fp16 activations; fp16 / bf16, uint4 g128 and
fp8-e4m3 128x128-block weights), enough for the reference's own `test_smoke_linear_correctness` to drive THESE GEMMs
unchanged.  Put this directory on PYTHONPATH (`import _turbomind` then resolves here); MoE entry points raise
NotImplementedError (grouped GEMM is outside the hot-path scope).

Tensors: a thin wrapper around a torch tensor (torch is the device-memory plumbing of this repository); DLPack goes
straight through it.  Quantisers restate the reference kernels' arithmetic with torch ops -- they produce TEST DATA and are
not on any hot path:
  QuantizeSymmBlock / DequantizeSymmBlock   kernels/quantization.cu:150-330   (scale = max(amax, 1e-8) / 448 per 128 x 128
                                            block in fp32, q = e4m3(float(x) * float(T(448 / amax))), x' = T(float(q) * scale))
  QuantizeGroupwise                         kernels/quantization.cu:423-436   (AWQ-style asymmetric u4 per group)
"""
import contextlib
import ctypes as C
import enum

import torch

from lmdeploy_b200 import _lib as L


class DataType(enum.Enum):
    TYPE_INVALID = 0
    TYPE_BOOL = 1
    TYPE_UINT8 = 2
    TYPE_INT32 = 6
    TYPE_FP16 = 10
    TYPE_FP32 = 11
    TYPE_BF16 = 14
    TYPE_FP8_E4M3 = 15
    TYPE_UINT4 = 17
    TYPE_FP4_E2M1 = 21


class Epilogue(enum.Enum):
    kNone = 0
    kGatedSilu = 0x2


_TORCH = {DataType.TYPE_FP16: torch.float16, DataType.TYPE_BF16: torch.bfloat16, DataType.TYPE_FP32: torch.float32,
          DataType.TYPE_UINT8: torch.uint8, DataType.TYPE_INT32: torch.int32, DataType.TYPE_FP8_E4M3: torch.float8_e4m3fn}
_FROM_TORCH = {v: k for k, v in _TORCH.items()}


class Tensor:
    """core::Tensor as Python sees it (bind.cpp:661-740): shape / type / DLPack / transposed view."""

    def __init__(self, t, dtype=None):
        self._t = t
        self._dtype = dtype or _FROM_TORCH.get(t.dtype, DataType.TYPE_INVALID)

    def __bool__(self):
        return self._t is not None and self._t.numel() > 0

    @property
    def shape(self):
        return list(self._t.shape)

    @property
    def type(self):
        return self._dtype

    @property
    def where(self):
        return "cuda" if self._t.is_cuda else "cpu"

    def byte_size(self):
        return self._t.numel() * self._t.element_size()

    def t(self):
        return Tensor(self._t.t(), self._dtype)

    def __dlpack__(self, stream=None):
        return self._t.__dlpack__() if stream is None else self._t.__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()

    def copy_from(self, obj):
        src = obj._t if isinstance(obj, Tensor) else torch.from_dlpack(obj)
        if src.numel() * src.element_size() != self.byte_size():
            raise RuntimeError("copy_from: byte size mismatch")          # bind.cpp:678-679
        self._t.view(torch.uint8).reshape(-1).copy_(src.contiguous().view(torch.uint8).reshape(-1))


def from_dlpack(obj):
    return obj if isinstance(obj, Tensor) else Tensor(torch.from_dlpack(obj))


from_dlpack_with_strides = from_dlpack


def generic_copy_on_stream(src, dst, stream_ptr):
    with torch.cuda.stream(torch.cuda.ExternalStream(stream_ptr)):
        dst._t.view(torch.uint8).reshape(-1).copy_(src._t.contiguous().view(torch.uint8).reshape(-1), non_blocking=True)


# ---------------------------------------------------------------------------------------------------------- configs
class _Format:
    def __init__(self, dtype, block_in=1, block_out=1):
        self.dtype, self.block_in, self.block_out = dtype, block_in, block_out


def ResolveLinearWeightFormat(data_type, weight_dtype, block_in, block_out):
    """core/data_format.cc:22-62: u4 -> {g, 1} with scales / zeros in the compute dtype; fp8 -> {128, 128}, f32 scales."""
    return _Format(weight_dtype, block_in, block_out)


class LinearConfig:
    def __init__(self):
        self.input_dim = self.output_dim = 0
        self.data_type = DataType.TYPE_FP16
        self.format = None
        self.has_bias = False


class _Param:
    def __init__(self):
        self._tensor = None

    def alloc(self, shape, dtype):
        if dtype == DataType.TYPE_UINT4:            # logical uint4 [K, N] stored two per byte (row-major, low nibble first)
            t = torch.zeros(shape[0], shape[1] // 2, dtype=torch.uint8, device="cuda")
        else:
            t = torch.zeros(*shape, dtype=_TORCH[dtype], device="cuda")
        self._tensor = Tensor(t, dtype)
        return self._tensor

    def get(self):
        return self._tensor

    def set(self, tensor):
        self._tensor = tensor


class _Desc:
    ld = 0
    offsets = 0


class LinearWeight:
    """models/linear_weight.h:42-112 over b200_linear_prepare / b200_linear_forward."""

    def __init__(self, cfg):
        self.input_dim, self.output_dim = cfg.input_dim, cfg.output_dim
        self.data_type = cfg.data_type
        self.weight_format = cfg.format or _Format(cfg.data_type)
        # no sm90 fp8 activation path here: activations stay in the compute dtype (linear_weight.cc:138-150)
        self.input_format = _Format(cfg.data_type)
        self.output_format = _Format(cfg.data_type)
        self.epilogue = Epilogue.kNone
        self.k_desc, self.q_desc = _Desc(), _Desc()
        self._params = {"weight": _Param(), "scales": _Param(), "zeros": _Param(), "bias": _Param()}
        self._desc = self._storage = self._ws = None

    def param(self, name):
        return self._params[name]

    def set_grouped(self, grouped):
        if grouped:
            raise NotImplementedError("grouped (MoE) weights are outside the hot-path scope")

    def set_fp8_fused_silu_output(self):
        raise NotImplementedError("the sm90 fp8 fused-SiLU output path does not exist on this architecture")

    def output_dtype(self):
        return self.data_type

    def prepare(self, gated=False):
        """LinearWeight::prepare (models/linear_weight.cc:101-324): loader-visible layout -> kernel layout.
        gated: the harness sets Epilogue.kGatedSilu AFTER prepare() on a weight whose columns are packed
        [gate block | up block] in 64-wide blocks (the sm90 layout); this library's gated epilogue pairs element-interleaved
        (gate, up) columns, so forward_dense re-runs prepare with the columns permuted (dense weights only)."""
        lib = L.load()
        self._gated = gated
        wt = self.weight_format.dtype
        dt = 0 if self.data_type == DataType.TYPE_FP16 else 1
        K, N = self.input_dim, self.output_dim
        w = self._params["weight"].get()._t
        st = torch.cuda.current_stream().cuda_stream
        if wt == DataType.TYPE_UINT4:
            if self.weight_format.block_in != 128:
                raise NotImplementedError("uint4 group size must be 128")
            fmt = L.FMT_U4
            qw = w.contiguous().view(torch.int32)                                     # [K, N / 8], nibble i = column 8w + i
            sc, zs = self._params["scales"].get()._t.contiguous(), self._params["zeros"].get()._t.contiguous()
            args = (qw.data_ptr(), sc.data_ptr(), zs.data_ptr())
            keep = (qw, sc, zs)
        elif wt == DataType.TYPE_FP8_E4M3:
            fmt = 2                                                                   # e4m3 + f32 scale per 128 x 128 block
            w8 = w.contiguous().view(torch.uint8)
            sc = self._params["scales"].get()._t.contiguous().float()
            args = (w8.data_ptr(), sc.data_ptr(), None)
            keep = (w8, sc)
        elif wt == self.data_type:
            fmt = L.FMT_DENSE
            wd = w.contiguous()
            if gated:
                blk = 64
                wd = wd.view(K, N // (2 * blk), 2, blk).transpose(2, 3).reshape(K, N).contiguous()
            args = (wd.data_ptr(), None, None)
            keep = (wd,)
        else:
            raise NotImplementedError(f"weight type {wt}")
        if gated and fmt != L.FMT_DENSE:
            raise NotImplementedError("block-packed fused SiLU exists for dense weights only here (the quantised form is the "
                                      "reference's sm90 path: per-block scales cannot follow an element-interleaved pairing)")
        self._desc = L.LinearDesc(K, N, fmt, dt, 128, None, None)
        nbytes = lib.b200_linear_prepared_bytes(C.byref(self._desc))
        if nbytes == 0:
            raise RuntimeError(lib.b200_last_error().decode())
        self._storage = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        self._desc.prepared = self._storage.data_ptr()
        L.check(lib.b200_linear_prepare(C.byref(self._desc), args[0], args[1], args[2], 0, self._storage.data_ptr(), st),
                "b200_linear_prepare")
        self._ws = torch.zeros(lib.b200_linear_workspace_bytes(C.byref(self._desc), 128), dtype=torch.uint8, device="cuda")
        del keep


class LlamaLinear:
    """models/llama/LlamaLinear.cu:226-273 (Forward) for dense weights; rows beyond 128 go through in chunks of 128 (the
    decode kernel's row capacity)."""

    def forward_dense(self, input, weight, output=None, input_scales=None, output_scales=None):
        if weight._desc is None:
            raise RuntimeError("LinearWeight.prepare() has not run")
        if input_scales:
            raise NotImplementedError("pre-quantised activations (sm90 fp8 path)")
        lib = L.load()
        x = input._t
        if x.stride(-1) != 1:
            x = x.contiguous()
        m, K = x.shape
        gated = weight.epilogue == Epilogue.kGatedSilu
        if gated and not getattr(weight, "_gated", False):
            weight.prepare(gated=True)
        nout = weight.output_dim // 2 if gated else weight.output_dim
        y = output._t if output else torch.empty(m, nout, dtype=x.dtype, device=x.device)
        st = torch.cuda.current_stream().cuda_stream
        for r0 in range(0, m, 128):
            rows = min(128, m - r0)
            L.check(lib.b200_linear_forward(C.byref(weight._desc), x[r0:].data_ptr(), x.stride(0), rows, y[r0:].data_ptr(),
                                            y.stride(0), 1 if gated else 0, weight._ws.data_ptr(), weight._ws.numel(), st),
                    "b200_linear_forward")
        return Tensor(y), None

    def forward_moe(self, *a, **k):
        raise NotImplementedError("MoE grouped GEMM is outside the hot-path scope")

    def set_measure(self, on):
        pass

    def import_records(self, path):
        return 0

    def export_records(self, path):
        return 0


# ---------------------------------------------------------------------------------------------------------- quantisers
def QuantizeSymmBlock(out=None, scale=None, src=None):
    x = src._t
    K, N = x.shape
    T = x.dtype
    xb = x.float().reshape(K // 128, 128, N // 128, 128)
    amax = xb.abs().amax(dim=(1, 3)).clamp_min(1e-8)
    s = (amax / 448.0).float()
    inv = (448.0 / amax).to(T).float()                                   # shared_inv_scale is T in the reference kernel
    q = (xb * inv[:, None, :, None]).reshape(K, N).to(torch.float8_e4m3fn)
    if out:
        out._t.view(torch.uint8).copy_(q.view(torch.uint8))
    else:
        out = Tensor(q, DataType.TYPE_FP8_E4M3)
    if scale:
        scale._t.copy_(s)
    else:
        scale = Tensor(s, DataType.TYPE_FP32)
    return out, scale


def DequantizeSymmBlock(out=None, src=None, scale=None):
    q = src._t.view(torch.float8_e4m3fn)
    K, N = q.shape
    T = out._t.dtype if out else torch.bfloat16
    y = (q.float().reshape(K // 128, 128, N // 128, 128) * scale._t.float()[:, None, :, None]).reshape(K, N).to(T)
    if out:
        out._t.copy_(y)
        return out
    return Tensor(y)


def QuantizeSymm(out=None, scale=None, src=None):
    raise NotImplementedError("activation quantisation belongs to the sm90 fp8 path")


def DequantizeSymm(out=None, src=None, scale=None):
    raise NotImplementedError("activation quantisation belongs to the sm90 fp8 path")


def QuantizeGroupwise(quant=None, scales=None, zeros=None, dequant=None, src=None, rbits=None, group_size=128):
    """Arguments are K-major views (N, K) as the reference's harness passes them (`.t()` of the [K, N] parameters)."""
    w = src._t.float()                                                    # [N, K]
    N, K = w.shape
    g = w.reshape(N, K // group_size, group_size)
    mn, mx = g.amin(-1), g.amax(-1)
    s = ((mx - mn).clamp_min(1e-5) / 15.0)
    z = (-torch.round(mn / s)).clamp(0, 15)
    q = (torch.round(g / s[..., None]) + z[..., None]).clamp(0, 15)
    T = scales._t.dtype
    sT, zT = s.to(T), z.to(T)
    dq = ((q - zT.float()[..., None]) * sT.float()[..., None]).reshape(N, K)
    scales._t.copy_(sT)                                                    # [N, K / g] view of the [K / g, N] parameter
    if zeros:
        zeros._t.copy_(zT)
    dequant._t.copy_(dq.to(dequant._t.dtype))
    qi = q.reshape(N, K).to(torch.uint8).t().contiguous()                 # [K, N]
    packed = (qi[:, 0::2] | (qi[:, 1::2] << 4)).contiguous()              # two columns per byte, low nibble first
    quant._t.t().copy_(packed)                                             # quant is the .t() view of the [K, N / 2] bytes


def MakeStridedPtrs(ptrs, dtype):
    raise NotImplementedError("MoE pointer tables are outside the hot-path scope")


def invokeMoeDispatch(**kw):
    raise NotImplementedError("MoE is outside the hot-path scope")


def invokeMoeCombine(**kw):
    raise NotImplementedError("MoE is outside the hot-path scope")


# ---------------------------------------------------------------------------------------------------------- context
class _DeviceContext:
    """core::ContextGuard as the harness uses it: a stream of its own, exposed as `stream_ptr`."""

    def __enter__(self):
        self._stream = torch.cuda.Stream()
        self.stream_ptr = self._stream.cuda_stream
        return self

    def __exit__(self, *exc):
        self._stream.synchronize()
        return False


def create_device_context():
    return _DeviceContext()
