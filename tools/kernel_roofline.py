"""Per-kernel roofline table of a decode step: algorithmic bytes / measured HBM peak -> ideal time, next to
the measured time.  Pure arithmetic (runs anywhere); measured times are passed in or default to the round-1 numbers.

    python tools/kernel_roofline.py [config=c2] [peak_gbs=6570.9] > profiles/<round>_kernel_roofline.md
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lmdeploy_b200.turbomind.specs import SPECS  # noqa: E402

CONFIGS = {"c2": ("internlm2-7b", 64, 2048, 8), "c3": ("llama3-8b", 128, 4096, 4), "c4": ("internlm2-20b", 64, 2048, 0)}
# measured on B200, round 1 (profiles/r01b_*.md; GEMMs graph-timed by tools/gemm_trace.cu, attention by bench.py)
MEASURED_US = {"c2": {"qkv": 12.8, "wo": 10.6, "w1w3": 22.6, "w2": 15.9, "attention": 56.6, "lm_head": 116.0,
                      "kv prologue": 5.7, "attention fix-up": 5.5, "residual+rmsnorm x2": 8.0}}


def weight_bytes(K, N, fmt):
    if fmt == "awq":
        return K * N // 2 + (K // 128) * N * 4
    if fmt == "fp8":
        return K * N + (K // 128) * N * 2
    return 2 * K * N


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6570.9
    name, bs, ctx, kvq = CONFIGS[cfg]
    s = SPECS[name]
    D = s.head_dim
    kv_bits = 16 if kvq == 0 else kvq
    per_tok = 2 * s.kv_heads * (D * kv_bits // 8 + (4 if kvq else 0))
    rows = [
        ("qkv", weight_bytes(s.hidden, (s.heads + 2 * s.kv_heads) * D, s.weight_format), s.layers),
        ("attention", bs * ctx * per_tok, s.layers),
        ("wo", weight_bytes(s.heads * D, s.hidden, s.weight_format), s.layers),
        ("w1w3", weight_bytes(s.hidden, 2 * s.inter, s.weight_format), s.layers),
        ("w2", weight_bytes(s.inter, s.hidden, s.weight_format), s.layers),
        ("lm_head", 2 * s.hidden * s.vocab, 1),
    ]
    meas = MEASURED_US.get(cfg, {})
    print(f"# Per-kernel roofline, {cfg}: {name}, bs {bs}, ctx {ctx}, KV {'fp16' if not kvq else 'int%d' % kvq}; "
          f"peak = {peak:.1f} GB/s (MEASURED_PEAKS.json)\n")
    print("| kernel | algorithmic bytes / launch | ideal µs | measured µs | fraction of peak | launches / step | ideal ms / step | measured ms / step |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    tot_i = tot_m = 0.0
    for k, nbytes, n in rows:
        ideal = nbytes / (peak * 1e9) * 1e6
        m = meas.get(k)
        tot_i += ideal * n / 1e3
        tot_m += (m or 0) * n / 1e3
        print(f"| {k} | {nbytes / 1e6:.1f} MB | {ideal:.1f} | {m if m is not None else '–'} | "
              f"{(ideal / m if m else float('nan')):.2f} | {n} | {ideal * n / 1e3:.3f} | {(m * n / 1e3 if m else float('nan')):.3f} |")
    glue = {k: v for k, v in meas.items() if k not in dict((r[0], 0) for r in rows)}
    for k, v in glue.items():
        tot_m += v * s.layers / 1e3
        print(f"| {k} | (latency-bound glue) | – | {v} | – | {s.layers} | – | {v * s.layers / 1e3:.3f} |")
    print(f"| **sum** | | | | | | **{tot_i:.3f}** | **{tot_m:.3f}** |")
    print(f"\nIdeal step at the measured peak: {tot_i:.3f} ms = {bs / tot_i * 1e3:,.0f} tokens/s; sum of measured kernel times: "
          f"{tot_m:.3f} ms.  `bench.py` (whole step, CUDA graph + PDL overlap): see profiles/r01b_bench.json.")
    print(json.dumps({"config": cfg, "ideal_ms": tot_i, "measured_kernel_sum_ms": tot_m}), file=sys.stderr)


if __name__ == "__main__":
    main()
