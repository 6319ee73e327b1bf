"""pipeline(): the user entry point, same shape as lmdeploy.pipeline (lmdeploy/api.py:15-82, pipeline.py:33-95)."""
from typing import List, Optional, Union

from .messages import GenerationConfig, Response, Session, TurbomindEngineConfig
from .turbomind.engine import TurboMind, batched_generate, stream_generate


class Pipeline:
    def __init__(self, model_path: str, backend_config: Optional[TurbomindEngineConfig] = None, tokenizer=None, **kw):
        self.backend_config = backend_config or TurbomindEngineConfig()
        self.engine = TurboMind.from_pretrained(model_path, self.backend_config, **kw)
        self.tokenizer = tokenizer
        self.hf_gen_cfg = {}
        import os
        if os.path.isdir(model_path):
            if tokenizer is None and os.path.exists(os.path.join(model_path, "tokenizer.json")):
                from transformers import AutoTokenizer
                self.tokenizer = AutoTokenizer.from_pretrained(model_path)
            # generation_config.json carries the chat models' real end-of-turn ids (Llama-3: [128001, 128009]); the
            # reference merges them into every request's stop ids (AsyncEngine.hf_gen_cfg, async_engine.py:417-437)
            gpath = os.path.join(model_path, "generation_config.json")
            if os.path.exists(gpath):
                import json
                with open(gpath) as fh:
                    self.hf_gen_cfg = json.load(fh)

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
        if len(cfgs) != len(ids):
            raise ValueError(f"{len(cfgs)} generation configs for {len(ids)} prompts")
        session_len = getattr(self.engine, "session_len", None) or self.engine.scheduler.session_len
        eos = getattr(self.tokenizer, "eos_token_id", None) if self.tokenizer is not None else self.engine.eos_id
        out = [(g or GenerationConfig()).determined(len(p), session_len, self.tokenizer, getattr(self, "hf_gen_cfg", None), eos)
               for g, p in zip(cfgs, ids)]
        return single, ids, out

    # ---- responses: what the reference's AsyncEngine.generate makes of the engine's output
    # (lmdeploy/serve/core/async_engine.py:656-790): the stop token that ended a sequence is NOT part of the response
    # (token_ids, text, generate_token_len) unless include_stop_str_in_output, in which case the closing chunk carries it;
    # a stream is content chunks with finish_reason None followed by ONE closing chunk with the finish_reason.
    def _decode(self, ids, g, special=None):
        if self.tokenizer is None or not ids:
            return ""
        skip = (g.skip_special_tokens if g is not None else True) if special is None else special
        return self.tokenizer.decode(list(ids), skip_special_tokens=skip)

    @staticmethod
    def _hidden_stop(s):
        """1 if the sequence ended on a stop token (it is then the last generated id), else 0."""
        return 1 if s.finished and s.finish_reason == "stop" and s.generated else 0

    def _closing(self, i, s, g, shown):
        """The closing Response of a prompt: finish_reason, and the stop token when the user asked to see it."""
        reason = {"cancel": "abort"}.get(s.finish_reason, s.finish_reason)          # the reference's vocabulary
        ids = []
        if self._hidden_stop(s) and g is not None and g.include_stop_str_in_output:
            ids = s.generated[-1:]
        return Response(text=self._decode(ids, g, special=False), generate_token_len=shown + len(ids),
                        input_token_len=len(s.prompt), finish_reason=reason, token_ids=list(ids), index=i,
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
        out = []
        for i, (s, g) in enumerate(zip(seqs, cfgs)):
            shown = s.generated[:len(s.generated) - self._hidden_stop(s)]
            r = self._closing(i, s, g, len(shown))
            r.token_ids = list(shown) + r.token_ids
            r.text = self._decode(shown, g) + r.text
            n = len(r.token_ids)           # entries of a hidden stop token are dropped with it (async_engine.py:741-748)
            if s.want_logprobs:
                r.logprobs = s.logprobs[:n]
            if s.want_logits and n:
                import torch
                r.logits = torch.stack(s.logits[:n])
            out.append(r)
        return out[0] if single else out

    __call__ = infer

    def batch_infer(self, *args, **kw):
        return self.infer(*args, **kw)

    def stream_infer(self, prompts, sessions=None, gen_config=None, do_preprocess: bool = True,
                     adapter_name: Optional[str] = None, stream_response: bool = True, **kw):
        """Generator of Response objects (lmdeploy/pipeline.py:145-178), `index` tells which prompt.  Content chunks
        (finish_reason None) carry the tokens and the text a prompt produced since its previous chunk -- every step with
        stream_response, once at the end without it -- and one closing chunk per prompt carries the finish_reason."""
        if adapter_name is not None or sessions is not None:
            raise NotImplementedError("sessions / adapters are outside the hot-path scope")
        _, ids, cfgs = self._prepare(prompts, gen_config)
        shown, text, held = {}, {}, {}
        for i, s, new in stream_generate(self.engine, ids, cfgs):
            g = cfgs[i]
            if s.finished and self._hidden_stop(s):
                new = new[:-1]
            held.setdefault(i, []).extend(new)
            shown[i] = shown.get(i, 0) + len(new)
            if held[i] and (stream_response or s.finished):
                # text delta of the whole visible sequence, so characters that span tokens come out whole; a trailing
                # replacement character means an unfinished byte sequence: hold the text back until it completes
                vis = s.generated[:shown[i]]
                full = self._decode(vis, g)
                prev = text.get(i, "")
                delta = full[len(prev):] if full.startswith(prev) else full
                if delta.endswith("\ufffd") and not s.finished:
                    delta = ""
                else:
                    text[i] = full
                r = Response(text=delta, generate_token_len=shown[i], input_token_len=len(s.prompt), finish_reason=None,
                             token_ids=list(held[i]), index=i, cached_tokens=s.reused_tokens)
                lo = shown[i] - len(held[i])
                if s.want_logprobs:
                    r.logprobs = s.logprobs[lo:shown[i]]
                if s.want_logits:
                    import torch
                    r.logits = torch.stack(s.logits[lo:shown[i]])
                yield r
                held[i] = []
            if s.finished:
                yield self._closing(i, s, g, shown.get(i, 0))

    # ---- multi-turn chat (lmdeploy/pipeline.py:186-266): the history is re-rendered through the chat template every
    # turn; with enable_prefix_caching the earlier turns' KV blocks are reused instead of recomputed
    def session(self) -> Session:
        self._next_chat_session = getattr(self, "_next_chat_session", 0) + 1
        return Session(self._next_chat_session - 1)

    @staticmethod
    def _history_to_messages(history, prompt):
        messages = []
        for user, assistant in history:
            messages += user if isinstance(user, list) else [{"role": "user", "content": user}]
            messages.append({"role": "assistant", "content": assistant})
        return messages + (prompt if isinstance(prompt, list) else [{"role": "user", "content": prompt}])

    def chat(self, prompt, session=None, gen_config: Optional[GenerationConfig] = None, stream_response=False,
             adapter_name=None, **kw):
        """One turn of a conversation: returns the updated Session (its .response holds the answer), or -- with
        stream_response -- an iterator of Response chunks that updates the session when exhausted."""
        if adapter_name is not None:
            raise NotImplementedError("LoRA adapters are outside the hot-path scope")
        if session is None:
            session = self.session()
        session.update(prompt=prompt, response=None)
        messages = self._history_to_messages(session.history, prompt)
        generator = self.stream_infer(messages, gen_config=gen_config, stream_response=stream_response)

        def _gen():
            resp = None
            for out in generator:
                # (a copy: the reference accumulates into the first chunk object it has already handed out)
                resp = resp.extend(out) if resp else Response(**{**vars(out), "token_ids": list(out.token_ids)})
                yield out
            session.response = resp
            session.history.append((session.prompt, resp.text))

        if stream_response:
            return _gen()
        for _ in _gen():
            pass
        return session

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
