"""tests/golden/ref_rope_configs.json: what the reference's loader hands its C++ RoPE config for a set of HF configs.

The reference functions (lmdeploy/turbomind/models/utils.py: parse_rope_param :76-166, copy_rope_config :170-187 and
their helpers) are EXECUTED from the reference source: the module cannot be imported here (it needs the `_turbomind`
extension), so the needed top-level definitions are picked out of its AST and run in a scratch namespace, then driven
exactly as make_attention_config drives them (:204-212).  Recorded per case: every field copy_rope_config writes into
the `_tm.RopeConfig` object, or the exception type.  tests/test_loader_cpu.py compares lmdeploy_b200's
parse_rope -> RopeSpec -> RopeConfig against it.

  python tools/make_rope_golden.py
"""
import ast
import json
import math
import os
from types import SimpleNamespace

REF = "/root/reference/lmdeploy/turbomind/models/utils.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_rope_configs.json")
WANTED = {"_optional_attr", "_param_get", "_param_has", "_ROPE_TYPE_MAP", "rope_type_to_int", "_get_mscale",
          "parse_rope_param", "copy_rope_config"}

BASE = {"max_position_embeddings": 131072, "rope_theta": 500000.0}
CASES = {
    "plain": (dict(BASE), 128),
    "theta_only_default_10000": ({"max_position_embeddings": 4096}, 128),
    "llama31": (dict(BASE, rope_scaling={"factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                                         "original_max_position_embeddings": 8192, "rope_type": "llama3"}), 128),
    "llama3_defaults": (dict(BASE, rope_scaling={"factor": 4.0, "rope_type": "llama3"}), 128),
    "linear_old_key": (dict(BASE, rope_scaling={"type": "linear", "factor": 2.0}), 128),
    "yarn_with_original": (dict(BASE, rope_scaling={"type": "yarn", "factor": 4.0,
                                                    "original_max_position_embeddings": 32768}), 128),
    "yarn_qwen25_style": ({"max_position_embeddings": 32768, "rope_theta": 1e6,
                           "rope_scaling": {"type": "yarn", "factor": 4.0, "original_max_position_embeddings": 32768}}, 128),
    "yarn_without_original": (dict(BASE, rope_scaling={"rope_type": "yarn", "factor": 2.0}), 128),
    "yarn_mscale": (dict(BASE, rope_scaling={"rope_type": "yarn", "factor": 40.0, "mscale": 1.0, "mscale_all_dim": 0.707,
                                             "beta_fast": 16, "beta_slow": 2}), 64),
    "yarn_attention_factor": (dict(BASE, rope_scaling={"rope_type": "yarn", "factor": 2.0, "attention_factor": 1.25}), 128),
    "partial_rotary": ({"max_position_embeddings": 4096, "partial_rotary_factor": 0.5}, 128),
    "rope_parameters_v5": ({"max_position_embeddings": 4096,
                            "rope_parameters": {"rope_theta": 1e6, "rope_type": "linear", "factor": 4.0,
                                                "partial_rotary_factor": 0.25}}, 128),
    "rope_parameters_default": ({"max_position_embeddings": 8192, "rope_parameters": {"rope_theta": 5e5, "rope_type": "default"}}, 128),
    "dynamic": (dict(BASE, rope_scaling={"rope_type": "dynamic", "factor": 2.0}), 128),
    "longrope": (dict(BASE, rope_scaling={"rope_type": "longrope"}), 128),
}


def load_reference_functions():
    tree = ast.parse(open(REF).read())
    keep = []
    for node in tree.body:
        name = getattr(node, "name", None)
        if name is None and isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
        if name in WANTED:
            keep.append(node)
    ns = {"math": math, "SimpleNamespace": SimpleNamespace, "__name__": "ref_utils_excerpt"}
    # `from __future__ import annotations` as in the module: its signatures use 3.9+ generics in annotations
    mod = ast.fix_missing_locations(ast.Module(body=keep, type_ignores=[]))
    exec(compile(mod, REF, "exec", flags=__import__("__future__").annotations.compiler_flag), ns)
    assert WANTED <= set(ns), WANTED - set(ns)
    return ns


def main():
    ns = load_reference_functions()
    out = {}
    for name, (cfg, head_dim) in CASES.items():
        try:
            rope, max_pos = ns["parse_rope_param"](cfg, head_dim)
            dst = SimpleNamespace()                      # stands for _tm.RopeConfig: records exactly what is assigned
            ns["copy_rope_config"](dst, rope, max_pos)
            out[name] = {"config": cfg, "head_dim": head_dim, "rope_config": vars(dst)}
        except Exception as e:
            out[name] = {"config": cfg, "head_dim": head_dim, "raises": type(e).__name__}
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, v.get("rope_config", v.get("raises")))


if __name__ == "__main__":
    main()
