"""ctypes binding of libb200tm.so.  The product path has NO fallback: if the CUDA library is missing
or cannot be loaded this raises immediately."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200_LIB_PATH: another build of the same library (A/B measurements of a kernel variant); still no fallback
LIB_PATH = os.environ.get("B200_LIB_PATH") or os.path.join(_HERE, "lib", "libb200tm.so")


class B200LibraryError(RuntimeError):
    pass


class RopeConfig(C.Structure):
    _fields_ = [
        ("type", C.c_int), ("dim", C.c_int), ("base", C.c_float), ("factor", C.c_float),
        ("max_position_embeddings", C.c_int),
        ("llama3_low_freq_factor", C.c_float), ("llama3_high_freq_factor", C.c_float),
        ("llama3_original_max_position_embeddings", C.c_int),
        ("yarn_attention_factor", C.c_float), ("yarn_beta_fast", C.c_float), ("yarn_beta_slow", C.c_float),
        ("exact_math", C.c_int),
    ]


class AttentionParams(C.Structure):
    _fields_ = [
        ("out", C.c_void_p), ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("stride", C.c_int64),
        ("q_bias", C.c_void_p), ("k_bias", C.c_void_p), ("v_bias", C.c_void_p),
        ("cu_q_len", C.c_void_p), ("cu_k_len", C.c_void_p), ("readonly_block_num", C.c_void_p),
        ("finished", C.c_void_p),
        ("block_ptrs", C.c_void_p), ("cu_block_nums", C.c_void_p), ("layer_offset", C.c_int),
        ("block_len", C.c_int),
        ("token_num", C.c_int), ("batch_size", C.c_int), ("max_k_len", C.c_int),
        ("num_heads", C.c_int), ("num_kv_heads", C.c_int), ("size_per_head", C.c_int),
        ("inv_sqrt_dh", C.c_float), ("window_size", C.c_int),
        ("rope", RopeConfig), ("quant_policy", C.c_int), ("dtype", C.c_int),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("plan", C.c_void_p),
        ("rope_table", C.c_void_p),
        ("out_packed_mt", C.c_int),
        ("debug_flags", C.c_int),
        ("stream", C.c_void_p),
    ]


class LinearDesc(C.Structure):
    _fields_ = [
        ("input_dim", C.c_int), ("output_dim", C.c_int), ("weight_format", C.c_int), ("dtype", C.c_int),
        ("group_size", C.c_int), ("prepared", C.c_void_p), ("bias", C.c_void_p),
    ]


class ModelConfig(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int), ("hidden", C.c_int), ("num_heads", C.c_int), ("num_kv_heads", C.c_int),
        ("head_dim", C.c_int), ("inter_size", C.c_int), ("vocab", C.c_int), ("embedding_rows", C.c_int),
        ("norm_eps", C.c_float), ("inv_sqrt_dh", C.c_float), ("rope", RopeConfig), ("dtype", C.c_int),
        ("quant_policy", C.c_int), ("block_len", C.c_int), ("max_batch", C.c_int), ("tp_size", C.c_int),
        ("tp_rank", C.c_int),
    ]


class SamplingParams(C.Structure):
    """b200_sampling_params (include/b200_turbomind.h)."""
    _fields_ = [
        ("logits", C.c_void_p), ("ld", C.c_int64), ("dtype", C.c_int), ("count", C.c_int), ("vocab", C.c_int),
        ("rows", C.c_void_p), ("temperature", C.c_void_p), ("top_k", C.c_void_p), ("top_p", C.c_void_p), ("min_p", C.c_void_p),
        ("repetition_penalty", C.c_void_p), ("history", C.c_void_p), ("history_len", C.c_void_p),
        ("rng_states", C.c_void_p), ("rng_index", C.c_void_p), ("out_ids", C.c_void_p), ("out_uniform", C.c_void_p),
        ("out_kept", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("stream", C.c_void_p),
    ]


class StepInputs(C.Structure):
    _fields_ = [
        ("batch_size", C.c_int), ("token_ids", C.c_void_p), ("cu_q_len", C.c_void_p), ("cu_k_len", C.c_void_p),
        ("finished", C.c_void_p), ("block_ptrs", C.c_void_p), ("cu_block_nums", C.c_void_p),
        ("layer_stride_bytes", C.c_int), ("max_k_len", C.c_int), ("out_tokens", C.c_void_p), ("advance", C.c_int),
    ]


ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_STATE = -1, -2, -3, -4          # b200_status
W_QKV, W_O, W_GATE_UP, W_DOWN, W_LM_HEAD = range(5)
FMT_DENSE, FMT_U4, FMT_FP8, FMT_FP8_COL = range(4)        # b200_linear_desc.weight_format
T_EMBEDDING, T_FINAL_NORM, T_ATTN_NORM, T_FFN_NORM, T_QKV_BIAS, T_Q_NORM, T_K_NORM = range(7)

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200LibraryError(
            f"{LIB_PATH} not found: build it with `python -m lmdeploy_b200.build` "
            "(lmdeploy_b200 has no CPU or PyTorch fallback)")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise B200LibraryError(f"cannot load {LIB_PATH}: {e}") from e
    lib.b200_last_error.restype = C.c_char_p
    lib.b200_version.restype = C.c_char_p
    lib.b200_decode_attention_workspace_bytes.restype = C.c_size_t
    lib.b200_decode_attention_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.b200_decode_attention_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.b200_decode_attention.argtypes = [C.POINTER(AttentionParams)]
    lib.b200_process_kv.argtypes = [C.POINTER(AttentionParams)]
    lib.b200_prefill_attention.argtypes = [C.POINTER(AttentionParams)]
    lib.b200_qk_rmsnorm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_float, C.c_int, C.c_void_p]
    lib.b200_rope_table.argtypes = [C.c_void_p, C.POINTER(RopeConfig), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.b200_prefill_attention_workspace_bytes.restype = C.c_size_t
    lib.b200_prefill_attention_workspace_bytes.argtypes = [C.c_int, C.c_int]
    lib.b200_decode_attention_grid.argtypes = [C.c_int, C.c_int]
    lib.b200_linear_prepared_bytes.restype = C.c_size_t
    lib.b200_linear_prepared_bytes.argtypes = [C.POINTER(LinearDesc)]
    lib.b200_linear_prepare.argtypes = [C.POINTER(LinearDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_void_p]
    lib.b200_linear_workspace_bytes.restype = C.c_size_t
    lib.b200_linear_workspace_bytes.argtypes = [C.POINTER(LinearDesc), C.c_int]
    lib.b200_linear_forward.argtypes = [C.POINTER(LinearDesc), C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                        C.c_int64, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.b200_linear_forward_packed.argtypes = [C.POINTER(LinearDesc), C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                               C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.b200_linear_forward_packed_residual_rmsnorm.argtypes = [
        C.POINTER(LinearDesc), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int,
        C.c_void_p, C.c_size_t, C.c_void_p]
    lib.b200_linear_plan.argtypes = [C.POINTER(LinearDesc), C.c_int, C.POINTER(C.c_int)]
    lib.b200_linear_plan.restype = C.c_int
    lib.b200_linear_load_plans.argtypes = [C.c_char_p]
    lib.b200_linear_load_plans.restype = C.c_int
    lib.b200_pack_activations.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p]
    lib.b200_rmsnorm.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int,
                                 C.c_void_p]
    lib.b200_residual_rmsnorm.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
    lib.b200_embedding_lookup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p]
    lib.b200_argmax.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p]
    lib.b200_sampling_rng_state_bytes.restype = C.c_size_t
    lib.b200_sampling_rng_state_bytes.argtypes = []
    lib.b200_sampling_workspace_bytes.restype = C.c_size_t
    lib.b200_sampling_workspace_bytes.argtypes = [C.c_int, C.c_int]
    lib.b200_sampling_init_rng.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.b200_sample.argtypes = [C.POINTER(SamplingParams)]
    lib.b200_comm_buffer_bytes.restype = C.c_size_t
    lib.b200_comm_buffer_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.b200_comm_flags_bytes.restype = C.c_size_t
    lib.b200_comm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p)]
    lib.b200_comm_destroy.argtypes = [C.c_void_p]
    lib.b200_comm_local_buffer.restype = C.c_void_p
    lib.b200_comm_local_buffer.argtypes = [C.c_void_p, C.c_int]
    lib.b200_comm_set_multicast.argtypes = [C.c_void_p, C.c_void_p]
    lib.b200_comm_set_geometry.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.b200_comm_set_geometry.restype = C.c_int
    lib.b200_comm_allreduce_residual_rmsnorm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                         C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                                         C.c_int, C.c_void_p]
    lib.b200_comm_argmax_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_void_p]
    lib.b200_comm_allreduce_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    lib.b200_model_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(ModelConfig)]
    lib.b200_model_destroy.argtypes = [C.c_void_p]
    lib.b200_model_set_linear.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(LinearDesc)]
    lib.b200_model_set_tensor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.b200_model_set_comm.argtypes = [C.c_void_p, C.c_void_p]
    lib.b200_model_buffer.restype = C.c_void_p
    lib.b200_model_buffer.argtypes = [C.c_void_p, C.c_int]
    lib.b200_model_decode_step.argtypes = [C.c_void_p, C.POINTER(StepInputs), C.c_void_p]
    lib.b200_model_chunk_step.argtypes = [C.c_void_p, C.POINTER(StepInputs), C.c_int, C.c_void_p]
    lib.b200_step_advance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.b200_ipc_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
    lib.b200_ipc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.b200_ipc_close.argtypes = [C.c_void_p]
    lib.b200_ipc_free.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().b200_last_error().decode()
        raise B200LibraryError(f"{what} failed (code {rc}): {msg}")
