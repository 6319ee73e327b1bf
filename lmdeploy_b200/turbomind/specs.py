"""Model geometries on the hot path (public HF configs of these models --
lmdeploy/turbomind/models/internlm2.py:70-84, models/llama.py:44-99 read the same fields)."""
from dataclasses import dataclass, replace


@dataclass(frozen=True)
class RopeSpec:
    """RoPE parameters as the reference parses them from an HF config (lmdeploy/turbomind/models/utils.py:76-166) and
    hands them to init_rope_kernel_param (src/turbomind/models/attention_weight.cc:37-93)."""
    type: str = "default"               # 'default' | 'linear' | 'llama3' | 'yarn' | 'dynamic' (prompts within the trained context)
    dim: int = 128                      # rotated dims = head_dim * partial_rotary_factor
    base: float = 10000.0
    factor: float = 0.0
    max_position_embeddings: int = 0    # the model's context length (yarn correction range; see loader.parse_rope)
    low_freq_factor: float = 1.0        # llama3
    high_freq_factor: float = 4.0
    original_max_position_embeddings: int = 8192
    attention_factor: float = 1.0       # yarn
    beta_fast: float = 32.0
    beta_slow: float = 1.0

    TYPE_ID = {"default": 1, "linear": 2, "yarn": 4, "llama3": 5}

    def dynamic_ntk_base(self, prompt_len):
        """The rope base the reference gives a request of `prompt_len` prompt tokens under dynamic NTK scaling
        (init_dynamic_ntk, src/turbomind/models/llama/unified_attention_layer.cc:228-243)."""
        base = self.base
        if self.type == "dynamic" and self.factor > 1.0 and prompt_len > self.max_position_embeddings:
            f = self.factor * prompt_len / self.max_position_embeddings - (self.factor - 1.0)
            base *= f ** (self.dim / (self.dim - 2.0))
        return base

    def config_fields(self):
        """Field values of b200_rope_config (include/b200_turbomind.h) in declaration order, without exact_math --
        the counterpart of copy_rope_config (lmdeploy/turbomind/models/utils.py:170-187)."""
        if self.type == "dynamic":
            # the kernels see plain RoPE with the config's base and no position scaling -- exactly what the reference
            # computes for a prompt that fits the trained context (RopeType::kDynamic: inv_factor = 1, base unchanged)
            return dict(RopeSpec(type="default", dim=self.dim, base=self.base, factor=0.0,
                                 max_position_embeddings=self.max_position_embeddings).config_fields())
        if self.type not in self.TYPE_ID:
            raise NotImplementedError(f"rope type {self.type!r} (supported: {sorted(self.TYPE_ID)})")
        return dict(type=self.TYPE_ID[self.type], dim=self.dim, base=self.base, factor=self.factor,
                    max_position_embeddings=self.max_position_embeddings,
                    llama3_low_freq_factor=self.low_freq_factor, llama3_high_freq_factor=self.high_freq_factor,
                    llama3_original_max_position_embeddings=self.original_max_position_embeddings,
                    yarn_attention_factor=self.attention_factor, yarn_beta_fast=self.beta_fast, yarn_beta_slow=self.beta_slow)


@dataclass(frozen=True)
class ModelSpec:
    name: str
    arch: str                 # 'internlm2' | 'llama' | 'qwen2'
    hidden: int
    layers: int
    heads: int
    kv_heads: int
    inter: int
    vocab: int
    rope_theta: float
    head_dim: int = 128
    norm_eps: float = 1e-5
    rope_type: str = "default"      # 'default' | 'llama3'
    rope_factor: float = 0.0
    weight_format: str = "awq"      # on-GPU weight format: 'awq' (u4 g128, any int4 checkpoint family) | 'fp16' | 'fp8'
    dtype: str = "f16"
    checkpoint_format: str = None   # how the checkpoint packs it: 'awq' | 'gptq' | 'compressed-tensors' | None
    qkv_bias: bool = False          # q/k/v projections carry a bias (Qwen2 family), applied by the attention prologue
    qk_norm: bool = False           # per-head RMSNorm of q and k before RoPE (Qwen3 family)
    rope: RopeSpec = None           # full RoPE description (checkpoints); None = rope_type / rope_theta / rope_factor above
    max_position_embeddings: int = 0   # config.max_position_embeddings (0 = unknown); feeds the RoPE config
    session_len: int = 0               # default session length the reference derives from the config (0 = unknown)

    def tiny(self, layers=2):
        return replace(self, layers=layers)


SPECS = {
    "internlm2-1.8b": ModelSpec("internlm2-1.8b", "internlm2", 2048, 24, 16, 8, 8192, 92544, 1e6, weight_format="fp16"),
    "internlm2-7b": ModelSpec("internlm2-7b", "internlm2", 4096, 32, 32, 8, 14336, 92544, 1e6),
    "internlm2-20b": ModelSpec("internlm2-20b", "internlm2", 6144, 48, 48, 8, 16384, 92544, 1e6),
    "llama3-8b": ModelSpec("llama3-8b", "llama", 4096, 32, 32, 8, 14336, 128256, 5e5, weight_format="fp16"),
    # Qwen2-7B: q/k/v bias, GQA group 7 (public HF config; lmdeploy/turbomind/models/qwen2.py)
    "qwen2-7b": ModelSpec("qwen2-7b", "qwen2", 3584, 28, 28, 4, 18944, 152064, 1e6, qkv_bias=True),
    "llama3-70b": ModelSpec("llama3-70b", "llama", 8192, 80, 64, 8, 28672, 128256, 5e5, weight_format="fp8",
                            dtype="bf16"),
}
