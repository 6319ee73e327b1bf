"""Public config / result types, mirroring lmdeploy/messages.py (field names and defaults of the reference:
TurbomindEngineConfig lmdeploy/messages.py:209-359, GenerationConfig :30-207, QuantPolicy :20-27, Response,
EngineOutput, ResponseType).  Only fields that influence the decode hot path are honoured; the others are accepted
and ignored so that user code written against the reference constructs them unchanged."""
import enum
from dataclasses import dataclass, field
from typing import List, Optional


class QuantPolicy(enum.IntEnum):
    NONE = 0
    INT4 = 4
    INT8 = 8


# max_batch_size left unset: the reference asks get_max_batch_size('cuda') (lmdeploy/utils.py:394-417), whose table has no
# entry for a B200 and answers 128 (tests/golden/ref_engine_defaults.json) -- also the row limit of the decode step
DEFAULT_MAX_BATCH_SIZE = 128
MAX_LOGPROBS = 1024            # lmdeploy/turbomind/turbomind.py MAX_LOGPROBS


class ConfigError(ValueError, AssertionError):
    """An invalid config value.  The reference validates with `assert` (lmdeploy/messages.py:195-205, 345-367), so code
    written against it may catch AssertionError; code written against this package catches ValueError.  Both work."""


class ResponseType(enum.Enum):
    """Same members as the reference (lmdeploy/messages.py ResponseType)."""
    SUCCESS = enum.auto()
    FINISH = enum.auto()
    ENGINE_STOP_ERROR = enum.auto()
    SESSION_REPEAT = enum.auto()
    SESSION_NOT_EXIST = enum.auto()
    HANDLER_NOT_EXIST = enum.auto()
    INPUT_LENGTH_ERROR = enum.auto()
    INTERNAL_ENGINE_ERROR = enum.auto()
    CANCEL = enum.auto()
    PREFIX_CACHE_CONFLICT = enum.auto()
    NO_QUEUE = enum.auto()
    NOT_SUPPORTED = enum.auto()
    OUT_OF_MEMORY = enum.auto()


@dataclass
class GenerationConfig:
    n: int = 1
    max_new_tokens: int = 512
    do_sample: bool = False
    top_p: float = 1.0
    top_k: int = 50
    min_p: float = 0.0
    temperature: float = 0.8
    repetition_penalty: float = 1.0
    ignore_eos: bool = False
    random_seed: Optional[int] = None
    stop_words: Optional[List[str]] = None
    bad_words: Optional[List[str]] = None
    stop_token_ids: Optional[List[int]] = None
    bad_token_ids: Optional[List[int]] = None
    min_new_tokens: Optional[int] = None
    skip_special_tokens: bool = True
    logprobs: Optional[int] = None
    output_logits: Optional[str] = None
    # accepted for keyword compatibility with the reference; anything but the default is refused (not on the decode
    # hot path: logits processors, guided decoding, hidden-state / ppl outputs, disaggregated serving, MoE routing)
    spaces_between_special_tokens: bool = True
    response_format: Optional[dict] = None
    logits_processors: Optional[list] = None
    output_last_hidden_state: Optional[str] = None
    include_stop_str_in_output: bool = False
    return_ppl: bool = False
    with_cache: bool = False
    preserve_cache: bool = False
    migration_request: Optional[object] = None
    return_routed_experts: bool = False
    repetition_ngram_size: int = 0
    repetition_ngram_threshold: int = 0

    _UNSUPPORTED = {"response_format": None, "logits_processors": None, "output_last_hidden_state": None, "return_ppl": False,
                    "with_cache": False, "preserve_cache": False, "migration_request": None, "return_routed_experts": False,
                    # acts on the logits before the argmax whatever do_sample says, so it cannot be ignored
                    "min_new_tokens": None}

    def convert_stop_bad_words_to_ids(self, tokenizer):
        """stop_words -> stop_token_ids with the reference's rule (lmdeploy/messages.py:154-174 and
        Tokenizer.indexes_containing_token, lmdeploy/tokenizer.py:152-190, 539-547): a word must encode to ONE token;
        every vocabulary entry whose token string contains it stops generation (' ' is looked up as the sentencepiece
        marker); more than 5 candidates are narrowed to the ids that decode to exactly the word.  `tokenizer` is a
        HuggingFace tokenizer (encode / decode / get_vocab)."""
        def ids_of(words):
            out = []
            for word in words or []:
                if not isinstance(word, str):
                    raise TypeError(f"stop_words must be a list of str but got {type(word)}")
                if len(tokenizer.encode(word, add_special_tokens=False)) > 1:
                    continue                       # multi-token words cannot be stop words (same as the reference)
                needle = "\u2581" if word == " " else word
                ids = [i for tok, i in tokenizer.get_vocab().items() if needle in tok]
                if len(ids) > 5:
                    ids = [i for i in sorted(ids) if tokenizer.decode([i]) == needle][:5]   # (the marker, as the reference)
                if not ids:
                    ids = tokenizer.encode(word, add_special_tokens=False)
                    ids = ids if len(ids) == 1 else []
                out += ids
            return out

        if self.bad_words or self.bad_token_ids:
            raise NotImplementedError("bad_words need a logits processor; the decode hot path samples greedily on the device")
        merged = set(ids_of(self.stop_words)) | set(self.stop_token_ids or [])
        self.stop_token_ids = sorted(merged) or None

    def update_from_hf_gen_cfg(self, generation_config, tokenizer_eos_token_id):
        """Add the tokenizer's and generation_config.json's eos ids to stop_token_ids (lmdeploy/messages.py:176-193)."""
        ids = set(self.stop_token_ids or [])
        if tokenizer_eos_token_id is not None:
            ids.add(int(tokenizer_eos_token_id))
        eos = (generation_config or {}).get("eos_token_id")
        if eos is not None:
            ids.update([eos] if isinstance(eos, int) else eos)
        self.stop_token_ids = sorted(ids)

    def __post_init__(self):
        # the reference's own checks (lmdeploy/messages.py:195-205)
        if type(self.n) is not int or self.n <= 0:
            raise ConfigError("n is not a positive integer")
        if not 0 <= self.top_p <= 1:
            raise ConfigError("top_p must be in [0, 1]")
        if self.top_k < 0:
            raise ConfigError("top_k can not be a negative integer")
        if not 0 <= self.temperature <= 2:
            raise ConfigError("temperature must be in [0, 2]")
        if not 0 <= self.min_p <= 1:
            raise ConfigError(f"min_p should be in range [0, 1], but found {self.min_p}")
        if self.repetition_ngram_size <= 0 or self.repetition_ngram_threshold <= 0:
            self.repetition_ngram_size = self.repetition_ngram_threshold = 0
        if self.max_new_tokens is not None and self.max_new_tokens < 0:
            raise ConfigError("max_new_tokens must be >= 0")
        # With do_sample=False -- the default -- the reference decodes greedily WHATEVER top_k / top_p / temperature /
        # repetition_penalty say (async_engine.py `_determine_gen_config` resets them).  do_sample=True runs the sampling
        # stage on the device (csrc/sampling.cu: repetition penalty, temperature, top-k <= 1024, top-p, min-p, per-request
        # random_seed); under tensor parallelism the logits are vocabulary-sharded and the engine refuses it at submission.
        if self.do_sample and self.top_k > 1024:
            raise ConfigError("top_k must be <= 1024 (the reference's top-k kernel limit, sampling_topk_kernels.cu:262-264)")
        # extra outputs read from the step's logits buffer (one GPU): the sampled row's logits / its top log-probabilities
        if self.output_logits not in (None, "generation"):
            raise NotImplementedError("output_logits='all' needs the logits of every prompt position; 'generation' is supported")
        if self.logprobs is not None:
            if self.logprobs < 1:
                raise ConfigError("logprobs must be a positive integer")
            self.logprobs = min(int(self.logprobs), MAX_LOGPROBS)       # the reference clips the same way (turbomind.py)
        if self.repetition_ngram_size:
            raise NotImplementedError("GenerationConfig.repetition_ngram_size is not supported on the greedy decode hot path")
        for name, default in self._UNSUPPORTED.items():
            if getattr(self, name) != default:
                raise NotImplementedError(f"GenerationConfig.{name} is not supported on the greedy decode hot path")

    def determined(self, input_len, session_len, tokenizer=None, hf_gen_cfg=None, eos_token_id=None):
        """The per-request copy the engine receives -- AsyncEngine._determine_gen_config
        (lmdeploy/serve/core/async_engine.py:417-437): stop words -> ids, eos ids of generation_config.json and of the
        tokenizer merged into stop_token_ids, greedy normalisation when do_sample is off, n > 1 falls back to 1,
        max_new_tokens None = whatever the session still holds."""
        import copy
        g = copy.copy(self)
        if g.stop_words or g.bad_words or g.bad_token_ids:
            if tokenizer is None and (g.stop_words or g.bad_words):
                raise ValueError("stop_words need a tokenizer: pass stop_token_ids instead")
            g.convert_stop_bad_words_to_ids(tokenizer)
        g.update_from_hf_gen_cfg(hf_gen_cfg or {}, eos_token_id)
        if not g.do_sample:
            g.top_k, g.temperature, g.repetition_penalty = 1, 1.0, 1.0
        g.n = 1
        if g.max_new_tokens is None:
            g.max_new_tokens = max(0, session_len - input_len)
        return g


@dataclass
class TurbomindEngineConfig:
    dtype: str = "auto"
    model_format: Optional[str] = None          # 'hf' | 'awq' | 'fp8'
    tp: int = 1
    dp: int = 1
    session_len: Optional[int] = None
    max_batch_size: Optional[int] = None
    cache_max_entry_count: float = 0.8          # ratio of free memory (<1) or block count (>=1)
    cache_chunk_size: int = -1
    cache_block_seq_len: int = 64
    enable_prefix_caching: bool = False
    quant_policy: int = 0                        # 0 | 4 | 8  (KV cache)
    rope_scaling_factor: float = 0.0
    use_logn_attn: bool = False
    download_dir: Optional[str] = None
    revision: Optional[str] = None
    max_prefill_token_num: int = 8192
    num_tokens_per_iter: int = 0
    max_prefill_iters: int = 1
    async_: int = 1
    devices: Optional[List[int]] = None
    empty_init: bool = False
    communicator: str = "native"                 # 'native' (NVLink peer kernels); 'nccl' is not on this path
    hf_overrides: Optional[dict] = None
    enable_metrics: bool = True
    # accepted for keyword compatibility with the reference (lmdeploy/messages.py TurbomindEngineConfig); the parallel /
    # multi-node / recurrent-cache knobs below must keep their defaults on this path
    cp: int = 1
    ep: int = 1
    device_num: Optional[int] = None
    attn_tp_size: Optional[int] = None
    attn_cp_size: Optional[int] = None
    attn_dp_size: Optional[int] = None
    mlp_tp_size: Optional[int] = None
    mlp_dp_size: Optional[int] = None
    outer_dp_size: Optional[int] = None
    nnodes: int = 1
    node_rank: int = 0
    dist_init_addr: Optional[str] = None
    cache_checkpoint_interval: int = 4096
    cache_prompt: str = "auto"
    cache_prompt_boundary_skip: int = 1
    cache_generation: str = "auto"
    language_model_only: bool = False

    def __post_init__(self):
        if self.dtype not in ("auto", "float16", "bfloat16"):
            raise ConfigError(f"unsupported dtype {self.dtype}")
        if self.tp < 1 or (self.tp & (self.tp - 1)):
            raise ConfigError("tp must be a power of two >= 1")
        if self.quant_policy not in (0, 4, 8):
            raise ConfigError("invalid quant_policy (0, 4, 8)")
        if self.cache_max_entry_count <= 0:
            raise ConfigError("invalid cache_max_entry_count")
        # the reference's remaining checks (lmdeploy/messages.py:345-367)
        if self.rope_scaling_factor < 0:
            raise ConfigError("invalid rope_scaling_factor")
        if self.max_prefill_token_num < 0:
            raise ConfigError("invalid max_prefill_token_num")
        if self.num_tokens_per_iter < 0:
            raise ConfigError("invalid num_tokens_per_iter")
        if self.cache_prompt not in ("all", "auto"):
            raise ConfigError("invalid cache_prompt")
        if self.cache_generation not in ("all", "auto", "none"):
            raise ConfigError("invalid cache_generation")
        if self.cache_checkpoint_interval <= 0:
            raise ConfigError("invalid cache_checkpoint_interval")
        if self.cache_prompt_boundary_skip < 1:
            raise ConfigError("invalid cache_prompt_boundary_skip")
        if self.async_ not in (0, 1):
            raise ConfigError("async_ must be 0 (disabled) or 1 (enabled)")
        # options with an effect this path does not implement: refused, not ignored
        if self.empty_init:
            raise NotImplementedError("empty_init (weights supplied later, e.g. by an RL trainer) is outside the hot-path scope")
        if self.devices is not None and list(self.devices) != list(range(self.tp)):
            raise NotImplementedError("devices: select GPUs with CUDA_VISIBLE_DEVICES; ranks use devices 0..tp-1")
        if self.cache_block_seq_len % 64:
            raise ValueError("cache_block_seq_len must be a multiple of 64")
        if self.max_batch_size is not None and not (1 <= self.max_batch_size <= 128):
            raise ValueError("max_batch_size must be in [1, 128] on the decode hot path")
        if self.dp != 1:
            raise NotImplementedError("data parallel = independent replicas; launch one engine per replica")
        if self.cp != 1 or self.ep != 1 or self.nnodes != 1 or (self.attn_cp_size or 1) != 1 or \
                (self.attn_dp_size or 1) != 1 or (self.mlp_dp_size or 1) != 1 or (self.outer_dp_size or 1) != 1:
            raise NotImplementedError("context / expert / data parallelism and multi-node launches are outside the hot path "
                                      "(tensor parallelism over one NVSwitch box only)")
        for name in ("attn_tp_size", "mlp_tp_size"):
            if getattr(self, name) not in (None, self.tp):
                raise NotImplementedError(f"{name} must equal tp")
        if self.communicator not in ("native", "nccl", "cuda-ipc"):
            raise ValueError(f"unknown communicator {self.communicator!r}")


@dataclass
class Response:
    """Fields of the reference's Response; finish_reason is one of 'stop' | 'length' | 'error' | 'abort'."""
    text: str = None
    generate_token_len: int = None
    input_token_len: int = None
    finish_reason: Optional[str] = None
    token_ids: List[int] = field(default_factory=list)
    logprobs: Optional[list] = None          # not produced on the greedy hot path
    logits: Optional[object] = None
    last_hidden_state: Optional[object] = None
    index: int = 0
    routed_experts: Optional[object] = None
    cached_tokens: int = 0                   # prompt tokens served from the prefix cache
    error_code: Optional[str] = None
    error_message: Optional[str] = None


    def extend(self, other: "Response") -> "Response":
        """Merge a later chunk into this one (lmdeploy/messages.py:634-660): text, token ids and logprobs are
        concatenated, everything else is taken from the later chunk.  Returns self."""
        self.text += other.text
        self.generate_token_len = other.generate_token_len
        self.input_token_len = other.input_token_len
        self.finish_reason = other.finish_reason
        self.index = other.index
        if other.token_ids:
            self.token_ids += other.token_ids
        if other.logprobs:
            self.logprobs = (self.logprobs or []) + other.logprobs
        self.routed_experts = other.routed_experts
        self.error_code = other.error_code
        self.error_message = other.error_message
        return self


@dataclass
class ScheduleMetrics:
    """The reference's ScheduleMetrics fields (lmdeploy/messages.py) plus the block counts they derive from."""
    active_seqs: int = 0
    waiting_seqs: int = 0
    cache_usage: float = 0.0
    prefix_cache_hit_rate: float = 0
    scheduler_tick: int = 0
    total_seqs: int = 0
    preemptions: int = 0
    total_blocks: int = 0
    active_blocks: int = 0
    cached_blocks: int = 0
    free_blocks: int = 0


class Session:
    """A chat session of Pipeline.chat: the turns so far (lmdeploy/serve/managers/session_manager.py:17-75)."""

    def __init__(self, session_id: int, **kwargs):
        self.session_id = session_id
        self.prompt = None
        self.response: Optional[Response] = None
        self.history: list = []                    # [(user prompt, assistant text)]
        self.gen_config: Optional[GenerationConfig] = None
        self.update(**kwargs)

    def update(self, **kwargs):
        self.prompt = kwargs.get("prompt", self.prompt)
        self.response = kwargs.get("response", self.response)
        self.gen_config = kwargs.get("gen_config", self.gen_config)

    def reset(self):
        self.prompt, self.response, self.history, self.gen_config = None, None, [], None

    def __repr__(self):
        return (f"Session(session_id={self.session_id}, history_len={len(self.history)}, "
                f"has_response={self.response is not None}, has_gen_config={self.gen_config is not None})")


@dataclass
class EngineOutput:
    status: ResponseType = None
    token_ids: List[int] = None
    num_token: int = 0
    logprobs: Optional[list] = None
    logits: Optional[object] = None
    last_hidden_state: Optional[object] = None
    cache_block_ids: Optional[List[int]] = None
    req_metrics: Optional[object] = None
    routed_experts: Optional[object] = None
    ce_loss: Optional[float] = None
