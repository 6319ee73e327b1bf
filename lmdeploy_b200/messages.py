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


class ResponseType(enum.Enum):
    SUCCESS = "success"
    FINISH = "finish"
    ENGINE_STOP_ERROR = "engine stop error"
    SESSION_REPEAT = "session repeat"
    SESSION_NOT_EXIST = "session not exist"
    INPUT_LENGTH_ERROR = "input length error"
    INTERNAL_ENGINE_ERROR = "internal engine error"
    CANCEL = "cancel"


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
                    ids = [i for i in sorted(ids) if tokenizer.decode([i]) == word][:5]
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
        if self.max_new_tokens is not None and self.max_new_tokens < 0:
            raise ValueError("max_new_tokens must be >= 0")
        # the decode hot path implements greedy sampling (top_k == 1, generation/sampling.cc:113-126)
        if self.do_sample and self.top_k != 1:
            raise NotImplementedError("lmdeploy_b200 implements the greedy decode path (top_k=1 / do_sample=False)")


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

    def __post_init__(self):
        if self.dtype not in ("auto", "float16", "bfloat16"):
            raise ValueError(f"unsupported dtype {self.dtype}")
        if self.tp < 1 or (self.tp & (self.tp - 1)):
            raise ValueError("tp must be a power of two >= 1")
        if self.quant_policy not in (0, 4, 8):
            raise ValueError("invalid quant_policy (0, 4, 8)")
        if self.cache_max_entry_count <= 0:
            raise ValueError("invalid cache_max_entry_count")
        if self.cache_block_seq_len % 64:
            raise ValueError("cache_block_seq_len must be a multiple of 64")
        if self.max_batch_size is not None and not (1 <= self.max_batch_size <= 128):
            raise ValueError("max_batch_size must be in [1, 128] on the decode hot path")
        if self.dp != 1:
            raise NotImplementedError("data parallel = independent replicas; launch one engine per replica")


@dataclass
class Response:
    text: str
    generate_token_len: int
    input_token_len: int
    finish_reason: Optional[str] = None
    token_ids: List[int] = field(default_factory=list)
    index: int = 0


@dataclass
class EngineOutput:
    status: ResponseType
    token_ids: List[int]
    num_token: int = 0
