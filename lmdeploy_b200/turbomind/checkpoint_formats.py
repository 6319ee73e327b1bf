"""Checkpoint tensor formats -> TurboMind orientation (the layout b200_linear_prepare consumes).

Every quantised HF checkpoint family the reference's TurboMind loader admits for int4 weights ends in the same
pre-prepare layout -- logical uint4 [K, N] + scales [K/g, N] + integer-valued zero points [K/g, N] -- and differs only
in how the checkpoint packs nibbles and which axis is major.  This module does that un-packing; the GPU never sees a
checkpoint-specific layout.  Behaviour follows lmdeploy/turbomind/weight_format.py:

  format               weight                         scales           zero points
  -------------------  -----------------------------  ---------------  --------------------------------------------
  awq ('gemm')         qweight int32 [K, N/8], nibble scales [K/g, N]  qzeros int32 [K/g, N/8], same nibble order
                       order [0,4,1,5,2,6,3,7] on N                    (:200-247)
  gptq (sym, no        qweight int32 [K/8, N], nibble scales [K/g, N]  qzeros int32 [K/g, N/8] plain order, stored
  act-order)           i = row 8r+i                                    minus one; absent -> 8              (:250-291)
  compressed-tensors   weight_packed int32 [N, K/8],  weight_scale     weight_zero_point int32 [N/8, K/g] packed on
  (pack-quantized)     plain order on K               [N, K/g]         N; absent -> 8                      (:294-346)
  fp8 (block 128x128)  weight e4m3 bytes [N, K]       weight_scale_inv f32 [N/128, K/128]                  (:349-393)
  dense                weight [N, K] float                                                                  (:177-197)

Selection is per linear and by what tensors exist at its prefix (the resolver of :431-502): a quantised checkpoint may
keep some linears dense (`modules_to_not_convert`), and those fall through to the dense format.  Pure tensor code --
checked on CPU against outputs of the reference module itself (tests/golden/ref_weight_formats.npz)."""
import torch

AWQ_NIBBLE_OF_COLUMN = (0, 4, 1, 5, 2, 6, 3, 7)     # column j of an 8-group lives in nibble AWQ_NIBBLE_OF_COLUMN[j]
PLAIN = tuple(range(8))
SYMMETRIC_ZERO = 8


def nibbles(x, order=PLAIN):
    """int32 [...] -> uint8 [..., 8]; element j of the new axis is nibble order[j] of the word."""
    if x.dtype != torch.int32:
        raise TypeError(f"packed int4 tensors are int32, got {x.dtype}")
    shifts = torch.tensor([4 * n for n in order], dtype=torch.int32)
    return ((x.unsqueeze(-1) >> shifts) & 0xF).to(torch.uint8)


def unpack_last(x, order=PLAIN):
    """[R, C/8] words packed along the last axis -> uint8 [R, C]."""
    return nibbles(x, order).reshape(*x.shape[:-1], x.shape[-1] * 8)


def unpack_first(x):
    """[R/8, C] words packed along the first axis (nibble i = row 8r+i) -> uint8 [R, C]."""
    return nibbles(x).permute(0, 2, 1).reshape(x.shape[0] * 8, x.shape[1])


def pack_u4_row(q):
    """uint8 [K, N] -> int32 [K, N/8], nibble i = column 8w+i   (pack_u4_row, weight_format.py:63-74)."""
    g = q.reshape(q.shape[0], -1, 8).to(torch.int64)
    w = torch.zeros(g.shape[:2], dtype=torch.int64)
    for i in range(8):
        w |= g[:, :, i] << (4 * i)
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)
    return w.to(torch.int32)


class Lin:
    """One linear in TurboMind orientation.  Exactly one payload is set:
         q uint8 [K,N] + scales f32 [K/g,N] + zeros f32 [K/g,N]      (int4, any checkpoint family)
         w f32 [K,N]                                                  (dense)
         w8 uint8 [K,N] + scales f32 [K/128, N]                       (fp8 e4m3; the checkpoint's 128x128 block scales
                                                                       expanded along N, as LinearWeight::prepare does
                                                                       before the reference kernel sees them,
                                                                       linear_weight.cc:138-150 -- so output columns can
                                                                       be permuted, interleaved and sliced freely)"""

    def __init__(self, q=None, scales=None, zeros=None, w=None, w8=None, bias=None):
        self.q, self.scales, self.zeros, self.w, self.w8, self.bias = q, scales, zeros, w, w8, bias

    @property
    def quant(self):
        return self.q is not None

    @property
    def fp8(self):
        return self.w8 is not None

    @property
    def kind(self):
        return "u4" if self.quant else "fp8" if self.fp8 else "dense"

    def cols(self, idx):
        b = None if self.bias is None else self.bias[idx]
        if self.quant:
            return Lin(self.q[:, idx], self.scales[:, idx], self.zeros[:, idx], bias=b)
        if self.fp8:
            return Lin(w8=self.w8[:, idx], scales=self.scales[:, idx], bias=b)
        return Lin(w=self.w[:, idx], bias=b)

    def rows(self, lo, hi, group=128):
        if self.quant or self.fp8:
            if lo % group or hi % group:
                raise ValueError(f"row slice [{lo}, {hi}) splits a quantisation group of {group}")
            sc = self.scales[lo // group:hi // group]
            if self.fp8:
                return Lin(w8=self.w8[lo:hi], scales=sc, bias=self.bias)
            return Lin(self.q[lo:hi], sc, self.zeros[lo // group:hi // group], bias=self.bias)
        return Lin(w=self.w[lo:hi], bias=self.bias)

    @staticmethod
    def cat(parts):
        kinds = {p.kind for p in parts}
        if len(kinds) != 1:
            raise NotImplementedError(f"fusing linears of different formats {sorted(kinds)} (the reference de-quantises "
                                      "mixed groups, linear.py dequant_mixed; not needed on the hot path)")
        biases = [p.bias for p in parts]
        if any(b is None for b in biases) != all(b is None for b in biases):
            raise ValueError("fusing linears with and without bias")
        b = None if biases[0] is None else torch.cat(biases)
        if parts[0].quant:
            return Lin(torch.cat([p.q for p in parts], 1), torch.cat([p.scales for p in parts], 1),
                       torch.cat([p.zeros for p in parts], 1), bias=b)
        if parts[0].fp8:
            return Lin(w8=torch.cat([p.w8 for p in parts], 1), scales=torch.cat([p.scales for p in parts], 1), bias=b)
        return Lin(w=torch.cat([p.w for p in parts], 1), bias=b)

    def dequant(self):
        """float32 [K, N] (host-side checks only)."""
        if self.quant:
            g = self.q.shape[0] // self.scales.shape[0]
            s = self.scales.repeat_interleave(g, 0)
            z = self.zeros.repeat_interleave(g, 0)
            return (self.q.to(torch.float32) - z) * s
        if self.fp8:
            s = self.scales.repeat_interleave(128, 0)[:self.w8.shape[0]]
            return self.w8.contiguous().view(torch.float8_e4m3fn).to(torch.float32) * s
        return self.w

    def to_tm(self, device, dtype):
        """-> the dict DecodeModel.load_tm_weights takes for one linear."""
        if self.quant:
            return dict(weight=pack_u4_row(self.q.contiguous()).to(device),
                        scales=self.scales.to(device, dtype).contiguous(),
                        zeros=self.zeros.to(device, dtype).contiguous())
        if self.fp8:
            return dict(weight=self.w8.contiguous().to(device), col_scales=self.scales.contiguous().to(device))
        return dict(weight=self.w.to(device, dtype).contiguous())


def _bias(av):
    b = av.get(".bias")
    return None if b is None else b.to(torch.float32)


class CheckpointFormat:
    name = ""
    suffixes = ()

    def __init__(self, group=128):
        self.group = group

    def accepts(self, av):
        raise NotImplementedError

    def read(self, av):
        raise NotImplementedError

    def _check_group(self, K, scales_rows, prefix_hint=""):
        if scales_rows * self.group != K:
            raise ValueError(f"{self.name}{prefix_hint}: {scales_rows} scale rows for K={K} is not group {self.group}")


class AwqFormat(CheckpointFormat):
    name = "awq"
    suffixes = (".qweight", ".scales", ".qzeros", ".bias")

    def accepts(self, av):
        qw, sc = av.get(".qweight"), av.get(".scales")
        if qw is None or qw.dtype != torch.int32:
            return False
        return sc is None or qw.shape[-1] * 8 == sc.shape[-1]

    def read(self, av):
        q = unpack_last(av[".qweight"], AWQ_NIBBLE_OF_COLUMN)
        scales = av[".scales"].to(torch.float16)
        self._check_group(q.shape[0], scales.shape[0])
        if ".qzeros" not in av:
            raise KeyError("awq linear without qzeros")
        zeros = unpack_last(av[".qzeros"], AWQ_NIBBLE_OF_COLUMN).to(torch.float16)
        return Lin(q, scales.to(torch.float32), zeros.to(torch.float32), bias=_bias(av))


class GptqFormat(CheckpointFormat):
    name = "gptq"
    suffixes = (".qweight", ".scales", ".qzeros", ".bias", ".g_idx")

    def accepts(self, av):
        qw, sc = av.get(".qweight"), av.get(".scales")
        if qw is None or qw.dtype != torch.int32:
            return False
        return sc is None or qw.shape[-1] == sc.shape[-1]

    def read(self, av):
        q = unpack_first(av[".qweight"])
        scales = av[".scales"].to(torch.float16)
        self._check_group(q.shape[0], scales.shape[0])
        gi = av.get(".g_idx")
        if gi is not None and not torch.equal(gi.to(torch.int64), torch.arange(q.shape[0]) // self.group):
            raise NotImplementedError("gptq act-order (desc_act) checkpoints are not supported (converter.py:183-185)")
        if ".qzeros" in av:
            zeros = (unpack_last(av[".qzeros"]) + 1).to(torch.float16)      # stored minus one
        else:
            zeros = torch.full(scales.shape, SYMMETRIC_ZERO, dtype=torch.float16)
        return Lin(q, scales.to(torch.float32), zeros.to(torch.float32), bias=_bias(av))


class CompressedTensorsFormat(CheckpointFormat):
    name = "compressed-tensors"
    suffixes = (".weight_packed", ".weight_scale", ".weight_zero_point", ".weight_shape", ".bias")

    def accepts(self, av):
        wp = av.get(".weight_packed")
        return wp is not None and wp.dtype == torch.int32

    def read(self, av):
        q = unpack_last(av[".weight_packed"]).t()                              # [N, K] -> [K, N]
        scales = av[".weight_scale"].to(torch.float16).t()
        self._check_group(q.shape[0], scales.shape[0])
        zp = av.get(".weight_zero_point")
        if zp is None:
            zeros = torch.full(scales.shape, SYMMETRIC_ZERO, dtype=torch.float16)
        elif zp.dtype == torch.int32:
            zeros = unpack_first(zp).to(torch.float16).t()                     # [N/8, K/g] -> [N, K/g] -> [K/g, N]
        else:
            zeros = zp.to(torch.float16).t()
        return Lin(q.contiguous(), scales.to(torch.float32).contiguous(), zeros.to(torch.float32).contiguous(),
                   bias=_bias(av))


class Fp8Format(CheckpointFormat):
    name = "fp8"
    suffixes = (".weight", ".weight_scale_inv", ".bias")

    def accepts(self, av):
        if ".weight_scale_inv" not in av:
            return False
        w = av.get(".weight")
        return w is None or w.dtype in (torch.float8_e4m3fn, torch.uint8)

    def read(self, av):
        w = av[".weight"]
        if w.dtype == torch.float8_e4m3fn:
            w = w.view(torch.uint8)
        w8 = w.t().contiguous()                                               # [N, K] -> [K, N]
        blocks = av[".weight_scale_inv"].to(torch.float32).t()               # [N/128, K/128] -> [K/128, N/128]
        if blocks.shape[0] * 128 < w8.shape[0] or blocks.shape[1] * 128 < w8.shape[1]:
            raise ValueError(f"fp8: {tuple(blocks.shape)} scale blocks do not cover a [{w8.shape[0]}, {w8.shape[1]}] weight")
        return Lin(w8=w8, scales=blocks.repeat_interleave(128, 1)[:, :w8.shape[1]].contiguous(), bias=_bias(av))


class DenseFormat(CheckpointFormat):
    name = "dense"
    suffixes = (".weight", ".bias")

    def accepts(self, av):
        if not set(av) <= {".weight", ".bias"}:
            return False
        w = av.get(".weight")
        return w is None or w.dtype.is_floating_point

    def read(self, av):
        return Lin(w=av[".weight"].to(torch.float32).t().contiguous(), bias=_bias(av))


_QUANT_FORMATS = {"awq": AwqFormat, "gptq": GptqFormat, "compressed-tensors": CompressedTensorsFormat,
                  "fp8": Fp8Format}
INT4_METHODS = ("awq", "gptq", "compressed-tensors")


class FormatResolver:
    """Ordered candidates: the checkpoint's quantised format (if any), then dense.  `read(tensors, prefix)` gathers the
    tensors that exist under `prefix` for the candidates' suffixes and hands them to the first format that accepts
    them.  Missing prefix -> KeyError; tensors no candidate accepts -> ValueError (weight_format.py:469-490)."""

    def __init__(self, quant_method=None, group=128):
        self.formats = []
        if quant_method not in (None, "hf", "fp16", "dense"):
            if quant_method not in _QUANT_FORMATS:
                raise ValueError(f"unknown checkpoint quantisation {quant_method!r}")
            self.formats.append(_QUANT_FORMATS[quant_method](group))
        self.formats.append(DenseFormat(group))
        self.suffixes = tuple(dict.fromkeys(s for f in self.formats for s in f.suffixes))

    def read(self, tensors, prefix):
        av = {s: tensors[prefix + s] for s in self.suffixes if prefix + s in tensors}
        if not av:
            raise KeyError(f"no checkpoint tensors at {prefix!r} (looked for {list(self.suffixes)})")
        for f in self.formats:
            if f.accepts(av):
                return f.read(av)
        raise ValueError(f"no weight format accepts the tensors at {prefix!r}: have {sorted(av)}, "
                         f"tried {[f.name for f in self.formats]}")


def parse_quantization_config(q):
    """HF `quantization_config` -> (method | None, group_size).  Admits what the reference admits for this path and
    what the sm_100a GEMM implements (int4 group 128; fp8 block 128): converter.py:95-110, 168-208."""
    if not q:
        return None, 0
    method = q.get("quant_method")
    if method == "awq":
        if q.get("version", "gemm").lower() != "gemm":
            raise NotImplementedError(f"awq version {q.get('version')!r}: only 'gemm' packing is supported")
        bits, group = q.get("bits", q.get("w_bit", 4)), q.get("group_size", q.get("q_group_size", 128))
    elif method == "gptq":
        if q.get("desc_act", False) or not q.get("sym", True):
            raise NotImplementedError("gptq checkpoints must be symmetric without act-order (desc_act=False, sym=True)")
        bits, group = q.get("bits", 4), q.get("group_size", 128)
    elif method == "compressed-tensors":
        grp = q["config_groups"]["group_0"]
        if grp.get("format", q.get("format")) != "pack-quantized":
            raise NotImplementedError("compressed-tensors: only the pack-quantized format is supported")
        w = grp["weights"]
        if w.get("type") != "int":
            raise NotImplementedError(f"compressed-tensors weights of type {w.get('type')!r}")
        bits, group = w.get("num_bits"), w.get("group_size")
    elif method == "fp8":
        blk = q.get("weight_block_size") or [128, 128]
        if list(blk) != [128, 128]:
            raise NotImplementedError(f"fp8 weight_block_size {blk}: only [128, 128]")
        return "fp8", 128
    else:
        raise NotImplementedError(f"quant_method {method!r} is outside the hot-path scope")
    if bits != 4:
        raise NotImplementedError(f"{method}: {bits}-bit weights (only 4-bit)")
    if group != 128:
        raise NotImplementedError(f"{method}: group_size {group} (the sm_100a W4A16 GEMM implements group 128)")
    return method, 128
