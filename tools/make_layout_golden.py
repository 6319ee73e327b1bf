"""tests/golden/ref_loader_layouts.json: the column order the reference's loader gives w_qkv and w1w3, per TP rank.

The pure tensor functions of the reference's loader path are EXECUTED from its source (the package cannot be imported
here: `_turbomind`, mmengine).  Their top-level definitions are taken from the AST, decorators dropped (the
`@transform_output_dim` wrapper only lifts a tensor function over the tensors of a Linear), and run on INDEX tensors --
row vectors whose value is the source column id -- in the order the builders call them:

  InternLM2 wqkv  -> _split_qkv_gqa                      lmdeploy/turbomind/models/internlm2.py:32-49
  q, k            -> _reorder_rotary_emb                  lmdeploy/turbomind/models/utils.py:306-324
  k, v            -> _repeat_kv_heads (tp)                lmdeploy/turbomind/builders/attention.py:30-42
  q, k, v         -> fuse_qkv (tp)                        lmdeploy/turbomind/builders/attention.py:65-76
  gate, up        -> _interleave_w1w3                     lmdeploy/turbomind/builders/ffn.py:31-34
  rank r          =  r-th contiguous chunk of the output dim (SplitSide.OUTPUT, builders/_base.py)

tests/test_loader_cpu.py runs lmdeploy_b200's convert_layer on checkpoints whose weights ARE column ids and compares.

  python tools/make_layout_golden.py
"""
import ast
import json
import os

import torch

REFDIR = "/root/reference/lmdeploy/turbomind"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_loader_layouts.json")
SOURCES = {"models/utils.py": ["_reorder_rotary_emb"], "builders/attention.py": ["_repeat_kv_heads", "fuse_qkv"],
           "builders/ffn.py": ["_interleave_w1w3"], "models/internlm2.py": ["_split_qkv_gqa"]}


def load_reference_functions():
    ns = {"torch": torch}
    for rel, names in SOURCES.items():
        path = os.path.join(REFDIR, rel)
        tree = ast.parse(open(path).read())
        keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
        assert len(keep) == len(names), (rel, names)
        for n in keep:
            n.decorator_list = []
        mod = ast.fix_missing_locations(ast.Module(body=keep, type_ignores=[]))
        exec(compile(mod, path, "exec", flags=__import__("__future__").annotations.compiler_flag), ns)
    return ns


def qkv_layout(ns, arch, Hq, Hkv, D, rope_dim, tp):
    """-> per-rank list of source column ids of w_qkv.  Source ids: HF order -- llama: [q | k | v] of the separate
    projections; internlm2: position inside the fused wqkv."""
    n = (Hq + 2 * Hkv) * D
    ids = torch.arange(n, dtype=torch.float32).view(1, n)
    if arch == "internlm2":
        q, k, v = ns["_split_qkv_gqa"](ids, head_dim=D, q_heads=Hq, kv_heads=Hkv)
    else:
        q, k, v = ids[:, :Hq * D], ids[:, Hq * D:(Hq + Hkv) * D], ids[:, (Hq + Hkv) * D:]
    q = ns["_reorder_rotary_emb"](q.contiguous(), D, rope_dim)
    k = ns["_reorder_rotary_emb"](k.contiguous(), D, rope_dim)
    k = ns["_repeat_kv_heads"](k, tp=tp, heads=Hkv)
    v = ns["_repeat_kv_heads"](v.contiguous(), tp=tp, heads=Hkv)
    merged = ns["fuse_qkv"](q, k, v, tp=tp)
    return [c.reshape(-1).to(torch.int64).tolist() for c in merged.chunk(tp, dim=-1)]


def ffn_layout(ns, inter, tp):
    """-> per-rank list of source column ids of w1w3; ids: gate = 0..inter-1, up = inter..2 inter-1."""
    g = torch.arange(inter, dtype=torch.float32).view(1, inter)
    fused = ns["_interleave_w1w3"](g, g + inter)
    return [c.reshape(-1).to(torch.int64).tolist() for c in fused.chunk(tp, dim=-1)]


def main():
    ns = load_reference_functions()
    out = {"qkv": [], "ffn": []}
    for arch in ("llama", "internlm2"):
        for (Hq, Hkv, D, rope_dim, tp) in [(4, 2, 128, 128, 1), (4, 2, 128, 128, 2), (4, 2, 128, 128, 4), (4, 2, 128, 64, 2),
                                           (8, 8, 128, 128, 4), (4, 1, 128, 128, 2)]:
            if arch == "internlm2" and rope_dim != D:
                continue      # the reference's own partial-rotary branch rejects the 3-D q _split_qkv_gqa returns
            out["qkv"].append(dict(arch=arch, heads=Hq, kv_heads=Hkv, head_dim=D, rope_dim=rope_dim, tp=tp,
                                   ranks=qkv_layout(ns, arch, Hq, Hkv, D, rope_dim, tp)))
    for inter, tp in [(512, 1), (512, 2), (512, 4)]:
        out["ffn"].append(dict(inter=inter, tp=tp, ranks=ffn_layout(ns, inter, tp)))
    json.dump(out, open(OUT, "w"), separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(out["qkv"]), "qkv cases,", len(out["ffn"]), "ffn cases")


if __name__ == "__main__":
    main()
