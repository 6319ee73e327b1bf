"""pipeline(): the user entry point, same shape as lmdeploy.pipeline (lmdeploy/api.py:15-82, pipeline.py:33-95)."""
from dataclasses import replace
from typing import List, Optional, Union

from .messages import GenerationConfig, Response, TurbomindEngineConfig
from .turbomind.engine import TurboMind, batched_generate


class Pipeline:
    def __init__(self, model_path: str, backend_config: Optional[TurbomindEngineConfig] = None, tokenizer=None, **kw):
        self.backend_config = backend_config or TurbomindEngineConfig()
        self.engine = TurboMind.from_pretrained(model_path, self.backend_config, **kw)
        self.tokenizer = tokenizer
        if tokenizer is None:
            import os
            if os.path.isdir(model_path) and os.path.exists(os.path.join(model_path, "tokenizer.json")):
                from transformers import AutoTokenizer
                self.tokenizer = AutoTokenizer.from_pretrained(model_path)

    def _encode(self, prompt):
        """A prompt is a list of token ids, a string, or OpenAI-style messages [{"role": ..., "content": ...}, ...]
        (rendered with the tokenizer's chat template, as the reference's pipeline does through its chat templates,
        lmdeploy/pipeline.py / lmdeploy/model.py)."""
        if isinstance(prompt, (list, tuple)) and prompt and isinstance(prompt[0], dict):
            if self.tokenizer is None:
                raise ValueError("chat messages need a tokenizer")
            ids = self.tokenizer.apply_chat_template(list(prompt), add_generation_prompt=True, tokenize=True)
            return list(ids["input_ids"] if isinstance(ids, dict) else ids)
        if isinstance(prompt, (list, tuple)):
            return [int(t) for t in prompt]
        if self.tokenizer is None:
            raise ValueError("no tokenizer: pass prompts as lists of token ids")
        return self.tokenizer.encode(prompt)

    def __call__(self, prompts: Union[str, List[int], List[str], List[List[int]]],
                 gen_config: Optional[GenerationConfig] = None, **kw):
        single = isinstance(prompts, str) or (isinstance(prompts, (list, tuple)) and prompts
                                              and isinstance(prompts[0], (int, dict)))
        plist = [prompts] if single else list(prompts)
        ids = [self._encode(p) for p in plist]
        if gen_config is not None and (gen_config.stop_words or gen_config.bad_words):
            if self.tokenizer is None:
                raise ValueError("stop_words need a tokenizer: pass stop_token_ids instead")
            gen_config = replace(gen_config)
            gen_config.convert_stop_bad_words_to_ids(self.tokenizer)
        seqs = batched_generate(self.engine, ids, gen_config)
        out = []
        for i, s in enumerate(seqs):
            text = self.tokenizer.decode(s.generated) if self.tokenizer is not None else ""
            reason = {"cancel": "abort"}.get(s.finish_reason, s.finish_reason)      # the reference's vocabulary
            out.append(Response(text=text, generate_token_len=len(s.generated), input_token_len=len(s.prompt),
                                finish_reason=reason, token_ids=list(s.generated), index=i,
                                cached_tokens=s.reused_tokens))
        return out[0] if single else out

    def close(self):
        self.engine = None


def pipeline(model_path: str, backend_config: Optional[TurbomindEngineConfig] = None, **kw) -> Pipeline:
    return Pipeline(model_path, backend_config, **kw)
