"""TEST INFRASTRUCTURE -- one TurboMind decode step assembled from the REFERENCE's OWN kernels compiled in place
(oracle/_ref: libtm_ref_attention.so, libtm_ref_gemm.so, libtm_ref_norm.so), run in the order the reference runs them:

    LanguageModel::Impl::Forward     models/language_model.cc:493-542   embedding rows -> decoder -> lm_head
    UnifiedDecoder::Forward          models/llama/unified_decoder.cc:226-346
        RMSNorm(layer 0) ; per layer: w_qkv GEMM -> dispatchDecoding -> wo GEMM -> residual+RMSNorm
                                      -> w1w3 GEMM (fused gated SiLU) -> w2 GEMM -> residual+RMSNorm(next norm)
    lm_head: dense T weights go to cuBLAS with fp32 compute (LlamaLinear.cu, cublas.cu:92-110) -> torch.matmul here
    greedy: argmax of the fp32-cast T logits (generation/generation.cc:270-276)

This is SURVEY section 8c's "L2" protocol: greedy token ids of the product against a forward built from reference kernels
on the same synthetic weights and KV pages.  Two honest deviations, both outside the kernels under test:
  * the step's first RMSNorm (no residual add) is run through the reference's fused residual+RMSNorm kernel with a zero
    addend (x + 0 == x in T), because invokeRMSNorm takes core::Tensor arguments the shim does not build;
  * embedding rows are gathered with torch indexing (the reference's lookup is a row copy).
Only tests/ may import this module."""
import ctypes as C

import numpy as np

from . import linear as OL
from . import ref_gemm_lib as RG
from . import ref_lib
from .numerics import F16, to_bits


def available():
    return ref_lib.available() and ref_lib.norm_available() and RG.available()


def _dev(x_f32, torch):
    """float32 numpy, exact in fp16 -> torch.float16 CUDA tensor with the same bits."""
    return torch.from_numpy(to_bits(np.ascontiguousarray(x_f32), F16).astype(np.int16)).cuda().view(torch.float16)


class ReferenceDecodeStep:
    """w: oracle.decode_step.TinyWeights(fmt='awq', dtype=F16).  All device buffers are torch tensors."""

    def __init__(self, w, torch, *, quant_policy, rope_config, inv_sqrt_dh, eps, block_len=64):
        if w.fmt != "awq" or w.dtype != F16:
            raise ValueError("the reference GEMM shim is built for u4 (g128) weights with fp16 activations")
        self.w, self.torch = w, torch
        self.quant_policy, self.rope, self.inv_sqrt_dh, self.eps, self.block_len = quant_policy, rope_config, inv_sqrt_dh, eps, block_len
        self.attn = ref_lib.load()
        self.norm = ref_lib.load_norm()
        self.gemm = RG.RefGemm()
        self.embed = _dev(w.embed, torch)
        self.final_norm = _dev(w.final_norm, torch)
        self.lm_head = _dev(w.output["w"], torch)          # [hidden, vocab]
        self.layers = []
        for lw in w.L:
            lin = {}
            for name, (K, N) in dict(w_qkv=(w.hidden, (w.heads + 2 * w.kv_heads) * w.D), wo=(w.heads * w.D, w.hidden),
                                     w1w3=(w.hidden, 2 * w.inter), w2=(w.inter, w.hidden)).items():
                d = lw[name]
                lin[name] = RG.RefLinear(self.gemm, torch.from_numpy(OL.pack_u4_row(d["q"])).cuda(), _dev(d["scales"], torch),
                                         _dev(d["zeros"], torch), K, N, torch)
            self.layers.append(dict(lin=lin, attn_norm=_dev(lw["attn_norm"], torch), ffn_norm=_dev(lw["ffn_norm"], torch)))
        self._keep = None

    def _residual_rmsnorm(self, residual, hidden, weight):
        """residual += hidden ; hidden <- RMSNorm(residual) * weight   (both in place, rms_norm.cu:287-363)"""
        rc = self.norm.ref_bias_residual_rmsnorm(residual.data_ptr(), hidden.data_ptr(), weight.data_ptr(), None,
                                                 residual.shape[1], residual.shape[0], self.eps, 0,
                                                 self.torch.cuda.current_stream().cuda_stream)
        if rc:
            raise RuntimeError(f"reference residual RMSNorm: CUDA error {rc}")

    def step(self, *, tokens, cu_q_len, cu_k_len, block_ptrs, cu_block_nums, max_k_len, layer_stride_bytes):
        """tokens int32 [B] (device).  Returns (logits fp16 [B, vocab], token ids int64 [B]), both on the device.  The
        new token's K/V are written into the pages behind block_ptrs exactly as the reference's decode prologue does."""
        from lmdeploy_b200 import _lib as L            # the argument block only (ref_lib translates it field by field)
        torch, w = self.torch, self.w
        B = tokens.shape[0]
        D, Hq, Hkv = w.D, w.heads, w.kv_heads
        qkv_dim = (Hq + 2 * Hkv) * D
        st = torch.cuda.current_stream().cuda_stream
        residual = self.embed[tokens.long()].contiguous()
        hidden = torch.zeros_like(residual)
        self._residual_rmsnorm(residual, hidden, self.layers[0]["attn_norm"])
        for l, lw in enumerate(self.layers):
            lin = lw["lin"]
            qkv = lin["w_qkv"].forward(hidden)
            out = torch.empty(B, Hq * D, dtype=torch.float16, device="cuda")
            p = L.AttentionParams()
            p.out, p.q = out.data_ptr(), qkv.data_ptr()
            p.k, p.v, p.stride = qkv.data_ptr() + Hq * D * 2, qkv.data_ptr() + (Hq + Hkv) * D * 2, qkv_dim
            p.cu_q_len, p.cu_k_len = cu_q_len.data_ptr(), cu_k_len.data_ptr()
            p.block_ptrs, p.cu_block_nums = block_ptrs.data_ptr(), cu_block_nums.data_ptr()
            p.layer_offset, p.block_len = l * layer_stride_bytes, self.block_len
            p.token_num, p.batch_size, p.max_k_len = B, B, max_k_len
            p.num_heads, p.num_kv_heads, p.size_per_head = Hq, Hkv, D
            p.inv_sqrt_dh, p.rope = self.inv_sqrt_dh, self.rope
            p.quant_policy, p.dtype, p.stream = self.quant_policy, 0, st
            a, self._keep = ref_lib.args_from_b200(p, torch, workspace=self._keep[:3] if self._keep else None)
            rc = self.attn.ref_decode_attention(C.byref(a))
            if rc:
                raise RuntimeError(f"reference dispatchDecoding: {rc}")
            o = lin["wo"].forward(out)
            self._residual_rmsnorm(residual, o, lw["ffn_norm"])
            act = lin["w1w3"].forward(o, gated_silu=True)
            y = lin["w2"].forward(act)
            nxt = self.layers[l + 1]["attn_norm"] if l + 1 < len(self.layers) else self.final_norm
            self._residual_rmsnorm(residual, y, nxt)
            hidden = y
        prev = torch.backends.cuda.matmul.allow_fp16_reduced_precision_reduction
        torch.backends.cuda.matmul.allow_fp16_reduced_precision_reduction = False       # cublasGemmEx, fp32 compute
        try:
            logits = hidden @ self.lm_head
        finally:
            torch.backends.cuda.matmul.allow_fp16_reduced_precision_reduction = prev
        return logits, logits.float().argmax(-1)

    def close(self):
        for lw in self.layers:
            for lin in lw["lin"].values():
                lin.close()
        self.gemm.close()
