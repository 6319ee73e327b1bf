"""tests/golden/ref_config_fields.json: field names and default expressions of the reference's public config dataclasses
(lmdeploy/messages.py: TurbomindEngineConfig, GenerationConfig), read from the SOURCE with ast -- the package itself cannot
be imported here (it needs mmengine).  tests/test_scheduler_cpu.py checks that lmdeploy_b200.messages accepts the same
keyword surface with the same defaults.

  python tools/make_config_golden.py
"""
import ast
import json
import os

REF = "/root/reference/lmdeploy/messages.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_config_fields.json")


def main():
    tree = ast.parse(open(REF).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("TurbomindEngineConfig", "GenerationConfig"):
            out[node.name] = {st.target.id: (ast.unparse(st.value) if st.value is not None else None)
                              for st in node.body if isinstance(st, ast.AnnAssign) and isinstance(st.target, ast.Name)}
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
