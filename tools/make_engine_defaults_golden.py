"""tests/golden/ref_engine_defaults.json: defaults the reference derives when the user leaves them unset.

  session_len     <- lmdeploy/utils.py:321-391 `_get_and_verify_max_len(hf_config, None)` (converter.py get_tm_config step 4):
                     the minimum over the context-length keys a config may carry, 2048 when it carries none
  max_batch_size  <- lmdeploy/utils.py:394-417 `get_max_batch_size('cuda')` for a device name that is not in its table
                     (a B200 is not): evaluated with torch.cuda.get_device_name stubbed to return the B200's name

Both functions are EXECUTED from the reference source (picked out of the AST; the package cannot be imported here).

  python tools/make_engine_defaults_golden.py
"""
import ast
import json
import os
import types

REF = "/root/reference/lmdeploy/utils.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_engine_defaults.json")

SESSION_CASES = {
    "llama3_8k": {"max_position_embeddings": 8192},
    "internlm2_32k": {"max_position_embeddings": 32768},
    "two_keys_min_wins": {"max_position_embeddings": 131072, "model_max_length": 4096},
    "seq_length_only": {"seq_length": 8192},
    "max_sequence_length_and_mpe": {"max_sequence_length": 2048, "max_position_embeddings": 4096},
    "nested_text_config": {"text_config": {"max_position_embeddings": 16384}, "max_position_embeddings": 99},
    "none": {"hidden_size": 4096},
}


class Cfg(dict):
    """dict with attribute access for nested configs, like a PretrainedConfig (hasattr + `in` both work)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return Cfg(v) if isinstance(v, dict) else v


def main():
    tree = ast.parse(open(REF).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("_get_and_verify_max_len", "get_max_batch_size")]
    assert len(keep) == 2
    quiet = types.SimpleNamespace(warning=lambda *a, **k: None, info=lambda *a, **k: None)
    ns = {"get_logger": lambda *a, **k: quiet}
    mod = ast.fix_missing_locations(ast.Module(body=keep, type_ignores=[]))
    exec(compile(mod, REF, "exec", flags=__import__("__future__").annotations.compiler_flag), ns)
    out = {"session_len": {k: {"config": c, "value": ns["_get_and_verify_max_len"](Cfg(c), None)} for k, c in SESSION_CASES.items()}}
    import torch
    real = torch.cuda.get_device_name
    try:
        sizes = {}
        for name in ("NVIDIA B200", "NVIDIA H100 80GB HBM3", "NVIDIA A100-SXM4-80GB"):
            torch.cuda.get_device_name = lambda *_a, _n=name: _n
            sizes[name] = ns["get_max_batch_size"]("cuda")
        out["max_batch_size"] = sizes
    finally:
        torch.cuda.get_device_name = real
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps({k: v["value"] for k, v in out["session_len"].items()}), out["max_batch_size"])


if __name__ == "__main__":
    main()
