"""pipeline(): the user entry point, same shape as lmdeploy.pipeline (lmdeploy/api.py:15-82, pipeline.py:33-95)."""
from dataclasses import replace
from typing import List, Optional, Union

from .messages import GenerationConfig, Response, TurbomindEngineConfig
from .turbomind.engine import TurboMind, batched_generate, stream_generate


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

    def _prepare(self, prompts, gen_config):
        single = isinstance(prompts, str) or (isinstance(prompts, (list, tuple)) and prompts
                                              and isinstance(prompts[0], (int, dict)))
        plist = [prompts] if single else list(prompts)
        ids = [self._encode(p) for p in plist]
        cfgs = gen_config if isinstance(gen_config, (list, tuple)) else [gen_config] * len(ids)
        out = []
        for g in cfgs:
            if g is not None and (g.stop_words or g.bad_words):
                if self.tokenizer is None:
                    raise ValueError("stop_words need a tokenizer: pass stop_token_ids instead")
                g = replace(g)
                g.convert_stop_bad_words_to_ids(self.tokenizer)
            out.append(g)
        return single, ids, out

    def _response(self, i, s, token_ids):
        reason = {"cancel": "abort"}.get(s.finish_reason, s.finish_reason)          # the reference's vocabulary
        text = self.tokenizer.decode(token_ids) if self.tokenizer is not None and token_ids else ""
        return Response(text=text, generate_token_len=len(s.generated), input_token_len=len(s.prompt),
                        finish_reason=reason if s.finished else None, token_ids=list(token_ids), index=i,
                        cached_tokens=s.reused_tokens, error_code="input_length" if s.error else None,
                        error_message=s.error)

    def infer(self, prompts: Union[str, List[int], List[str], List[List[int]], List[dict]],
              gen_config: Union[GenerationConfig, List[GenerationConfig], None] = None, do_preprocess: bool = True,
              adapter_name: Optional[str] = None, use_tqdm: bool = False, **kw):
        """Same call shape as the reference's Pipeline.infer (lmdeploy/pipeline.py:97-139): a prompt or a list of
        prompts (text, token ids or chat messages), one GenerationConfig or one per prompt -> Response / [Response]."""
        if adapter_name is not None:
            raise NotImplementedError("LoRA adapters are outside the hot-path scope")
        single, ids, cfgs = self._prepare(prompts, gen_config)
        seqs = batched_generate(self.engine, ids, cfgs)
        out = [self._response(i, s, s.generated) for i, s in enumerate(seqs)]
        return out[0] if single else out

    __call__ = infer

    def batch_infer(self, *args, **kw):
        return self.infer(*args, **kw)

    def stream_infer(self, prompts, sessions=None, gen_config=None, do_preprocess: bool = True,
                     adapter_name: Optional[str] = None, stream_response: bool = True, **kw):
        """Generator of Response objects (lmdeploy/pipeline.py:145-178): with stream_response every yield carries the tokens
        a prompt produced since its previous yield (`index` tells which prompt; finish_reason is set on its last one);
        without it one final Response per prompt."""
        if adapter_name is not None or sessions is not None:
            raise NotImplementedError("sessions / adapters are outside the hot-path scope")
        _, ids, cfgs = self._prepare(prompts, gen_config)
        acc = {}
        for i, s, new in stream_generate(self.engine, ids, cfgs):
            if stream_response:
                if new or s.finished:
                    yield self._response(i, s, new)
            else:
                acc.setdefault(i, []).extend(new)
                if s.finished:
                    yield self._response(i, s, acc[i])

    def get_ppl(self, input_ids):
        raise NotImplementedError("get_ppl needs per-position logits of the prompt (return_ppl path); the decode hot path "
                                  "keeps only the sampled row")

    def close(self):
        eng, self.engine = self.engine, None
        if eng is not None and hasattr(eng, "close"):
            eng.close()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.close()


def pipeline(model_path: str, backend_config: Optional[TurbomindEngineConfig] = None, chat_template_config=None,
             log_level: str = "WARNING", max_log_len: Optional[int] = None, trust_remote_code: bool = False,
             speculative_config=None, allowed_media_domains=None, **kw) -> Pipeline:
    """Same signature as lmdeploy.pipeline (lmdeploy/api.py:15-82).  chat_template_config: the tokenizer's own chat
    template is used; log_level / max_log_len / trust_remote_code / allowed_media_domains have nothing to act on here;
    speculative decoding is not on this path."""
    if speculative_config is not None:
        raise NotImplementedError("speculative decoding is outside the hot-path scope")
    return Pipeline(model_path, backend_config, **kw)
