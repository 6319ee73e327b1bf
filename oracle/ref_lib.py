"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_ref/libtm_ref_attention.so: the REFERENCE's own decode
attention + ProcessKV kernels (src/turbomind/kernels/attention/{decoding.cu,reduce.cu,kv_cache_utils_v2.cu,
kernel/decoding_sm80_128.cu}) compiled in place for sm_100a by oracle/ref_build/Makefile.

Only tests/, __graft_entry__.smoke() and bench.py's reference-kernel leg may import this module; the product
path (lmdeploy_b200/) never does.  The .so is built in the authoring container (where /root/reference exists)
and travels to the GPU box with the repo snapshot; nothing here reads /root/reference at run time.
"""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(ROOT, "_ref", "libtm_ref_attention.so")


class RefRope(C.Structure):
    _fields_ = [("type", C.c_int), ("dim", C.c_int), ("base", C.c_float), ("factor", C.c_float),
                ("max_position_embeddings", C.c_int), ("yarn_attention_factor", C.c_float),
                ("yarn_beta_fast", C.c_float), ("yarn_beta_slow", C.c_float),
                ("llama3_low_freq_factor", C.c_float), ("llama3_high_freq_factor", C.c_float),
                ("llama3_original_max_position_embeddings", C.c_int)]


class RefAttnArgs(C.Structure):
    _fields_ = [("out", C.c_void_p), ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
                ("stride", C.c_longlong),
                ("q_bias", C.c_void_p), ("k_bias", C.c_void_p), ("v_bias", C.c_void_p),
                ("cu_q_len", C.c_void_p), ("cu_k_len", C.c_void_p), ("readonly_block_num", C.c_void_p),
                ("finished", C.c_void_p), ("block_ptrs", C.c_void_p), ("cu_block_nums", C.c_void_p),
                ("layer_offset", C.c_int), ("block_len", C.c_int), ("token_num", C.c_int),
                ("batch_size", C.c_int), ("max_q_len", C.c_int), ("max_k_len", C.c_int),
                ("num_heads", C.c_int), ("num_kv_heads", C.c_int), ("size_per_head", C.c_int),
                ("inv_sqrt_dh", C.c_float), ("rope", RefRope), ("quant_policy", C.c_int),
                ("max_split_k", C.c_int), ("split_cnt", C.c_void_p), ("partial_O", C.c_void_p),
                ("partial_ML", C.c_void_p), ("dtype", C.c_int), ("stream", C.c_void_p)]


PREFILL_LIB_PATH = os.path.join(ROOT, "_ref", "libtm_ref_prefill.so")

_lib = None
_prefill_lib = None


def available():
    return os.path.exists(LIB_PATH)


NORM_LIB_PATH = os.path.join(ROOT, "_ref", "libtm_ref_norm.so")
_norm_lib = None


def norm_available():
    return os.path.exists(NORM_LIB_PATH)


def load_norm():
    """libtm_ref_norm.so: the reference's invokeBiasResidualRMSNorm (oracle/ref_build/ref_norm_shim.cu)."""
    global _norm_lib
    if _norm_lib is None:
        _norm_lib = C.CDLL(NORM_LIB_PATH)
        _norm_lib.ref_bias_residual_rmsnorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                        C.c_float, C.c_int, C.c_void_p]
        _norm_lib.ref_bias_residual_rmsnorm.restype = C.c_int
        if hasattr(_norm_lib, "ref_qk_rmsnorm"):          # invokeQkRMSNorm (Qwen3 family)
            _norm_lib.ref_qk_rmsnorm.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                 C.c_int, C.c_float, C.c_int, C.c_void_p]
            _norm_lib.ref_qk_rmsnorm.restype = C.c_int
    return _norm_lib


def prefill_available():
    return os.path.exists(PREFILL_LIB_PATH)


def load_prefill():
    """libtm_ref_prefill.so: the reference's ProcessKV + FlattenKV + dispatchAttention for a chunk
    (oracle/ref_build/ref_prefill_shim.cu)."""
    global _prefill_lib
    if _prefill_lib is None:
        _prefill_lib = C.CDLL(PREFILL_LIB_PATH)
        _prefill_lib.ref_prefill_attention.argtypes = [C.POINTER(RefAttnArgs), C.c_void_p, C.c_int]
        _prefill_lib.ref_prefill_attention.restype = C.c_int
    return _prefill_lib


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_decode_attention.argtypes = [C.POINTER(RefAttnArgs)]
        _lib.ref_decode_attention.restype = C.c_int
        _lib.ref_process_kv.argtypes = [C.POINTER(RefAttnArgs)]
        _lib.ref_process_kv.restype = C.c_int
        _lib.ref_version.restype = C.c_char_p
    return _lib


def args_from_b200(p, torch, *, max_q_len=1, workspace=None):
    """Translate a lmdeploy_b200._lib.AttentionParams into the reference's argument block (same pointers)."""
    a = RefAttnArgs()
    for f in ("out", "q", "k", "v", "stride", "q_bias", "k_bias", "v_bias", "cu_q_len", "cu_k_len",
              "readonly_block_num", "finished", "block_ptrs", "cu_block_nums", "layer_offset", "block_len",
              "token_num", "batch_size", "max_k_len", "num_heads", "num_kv_heads", "size_per_head", "inv_sqrt_dh",
              "quant_policy", "dtype", "stream"):
        setattr(a, f, getattr(p, f))
    a.max_q_len = max_q_len
    r = p.rope
    a.rope = RefRope(r.type, r.dim, r.base, r.factor, r.max_position_embeddings, r.yarn_attention_factor,
                     r.yarn_beta_fast, r.yarn_beta_slow, r.llama3_low_freq_factor, r.llama3_high_freq_factor,
                     r.llama3_original_max_position_embeddings)
    a.max_split_k = 0
    if workspace is None:
        H, D = p.num_heads, p.size_per_head
        workspace = (torch.zeros(4096, dtype=torch.int32, device="cuda"),
                     torch.zeros(4096 * H * D, dtype=torch.float32, device="cuda"),
                     torch.zeros(4096 * H * 2, dtype=torch.float32, device="cuda"))
    a.split_cnt, a.partial_O, a.partial_ML = (t.data_ptr() for t in workspace[:3])
    if not a.finished:
        # the reference kernel reads finished[batch] unconditionally (attention_universal.h:380)
        fin = torch.zeros(max(1, p.batch_size), dtype=torch.uint8, device="cuda")
        workspace = tuple(workspace) + (fin,)
        a.finished = fin.data_ptr()
    return a, workspace
