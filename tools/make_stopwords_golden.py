"""tests/golden/ref_stop_words.json: stop_words -> stop_token_ids as the REFERENCE computes them.

Executed from the reference source (AST excerpts; the package cannot be imported here):
  GenerationConfig.convert_stop_bad_words_to_ids / update_from_hf_gen_cfg   lmdeploy/messages.py:154-193
  Tokenizer.indexes_containing_token                                        lmdeploy/tokenizer.py:539-547
  HuggingFaceTokenizer.indexes_containing_token                             lmdeploy/tokenizer.py:152-190
over the toy vocabulary of tests/fake_tokenizer.py (no real tokenizer exists in this image).

  python tools/make_stopwords_golden.py
"""
import ast
import json
import os
import sys
import types
from collections import deque

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fake_tokenizer import FakeTokenizer  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_stop_words.json")
CASES = [
    dict(stop_words=["\n", "ld", " Hello wor"], stop_token_ids=[1]),
    dict(stop_words=["!"]),
    dict(stop_words=[" "]),
    dict(stop_words=["<|im_end|>"]),
    dict(stop_words=["x"]),
    dict(stop_words=None, stop_token_ids=[3, 3, 2]),
    dict(stop_words=[], stop_token_ids=None),
    dict(stop_words=["</s>", "\n\n"], stop_token_ids=[9]),
]
EOS_CASES = [
    dict(stop_token_ids=[5], generation_config={"eos_token_id": 2}, tokenizer_eos=1),
    dict(stop_token_ids=None, generation_config={"eos_token_id": [2, 7]}, tokenizer_eos=None),
    dict(stop_token_ids=[7], generation_config={}, tokenizer_eos=7),
]


def methods(path, cls, names):
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
    keep = [n for n in node.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(keep) == len(names), (cls, names)
    ns = {}
    mod = ast.fix_missing_locations(ast.Module(body=keep, type_ignores=[]))
    exec(compile(mod, path, "exec", flags=__import__("__future__").annotations.compiler_flag), ns)
    return ns


def reference_tokenizer(hf):
    """The reference's Tokenizer -> HuggingFaceTokenizer pair around `hf`, with exactly the attributes the two
    indexes_containing_token methods read (tokenizer.py:60-67 initialises them so)."""
    quiet = types.SimpleNamespace(warning=lambda *a, **k: None)
    inner_fn = methods("/root/reference/lmdeploy/tokenizer.py", "HuggingFaceTokenizer", ["indexes_containing_token"])
    outer_fn = methods("/root/reference/lmdeploy/tokenizer.py", "Tokenizer", ["indexes_containing_token"])
    inner = types.SimpleNamespace(model=hf, token2id={}, _indexes_tokens_deque=deque(maxlen=10), max_indexes_num=5,
                                  maybe_decode_bytes=False, vocab_size=len(hf), logger=quiet,
                                  decode=lambda ids: hf.decode(ids),
                                  encode=lambda s, add_bos=True, **kw: hf.encode(s, add_special_tokens=add_bos))
    inner.indexes_containing_token = lambda tok: inner_fn["indexes_containing_token"](inner, tok)
    outer = types.SimpleNamespace(model=inner, logger=quiet,
                                  encode=lambda s, add_bos=True, **kw: hf.encode(s, add_special_tokens=add_bos))
    outer.indexes_containing_token = lambda tok: outer_fn["indexes_containing_token"](outer, tok)
    return outer


def main():
    gen = methods("/root/reference/lmdeploy/messages.py", "GenerationConfig",
                  ["convert_stop_bad_words_to_ids", "update_from_hf_gen_cfg"])
    out = {"vocab": FakeTokenizer.VOCAB, "stop_words": [], "eos": []}
    for case in CASES:
        cfg = types.SimpleNamespace(stop_words=case.get("stop_words"), bad_words=None,
                                    stop_token_ids=case.get("stop_token_ids"), bad_token_ids=None)
        gen["convert_stop_bad_words_to_ids"](cfg, reference_tokenizer(FakeTokenizer()))
        out["stop_words"].append(dict(case, result=sorted(cfg.stop_token_ids) if cfg.stop_token_ids is not None else None))
    for case in EOS_CASES:
        cfg = types.SimpleNamespace(stop_token_ids=case["stop_token_ids"])
        gen["update_from_hf_gen_cfg"](cfg, case["generation_config"], case["tokenizer_eos"])
        out["eos"].append(dict(case, result=sorted(cfg.stop_token_ids)))
    json.dump(out, open(OUT, "w"), indent=1, ensure_ascii=True)
    for c in out["stop_words"] + out["eos"]:
        print(c)


if __name__ == "__main__":
    main()
