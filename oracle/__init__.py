"""CPU oracle for the TurboMind decode hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference algorithm (InternLM/lmdeploy
0.16.0, src/turbomind) for the decode hot path named in BASELINE.json.  Each
function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker / timed
CPU baseline.  Nothing under ``lmdeploy_b200/`` imports it; the product path
fails loudly when the CUDA library is missing.

Parity pinning: PINNED against the reference itself.  The reference stores no golden
vectors for this path and its build system cannot be configured
offline, but its decode-attention translation units compile from their sources in place:
oracle/ref_build/ builds them (unmodified, from /root/reference) into
oracle/_ref/libtm_ref_attention.so, and
  * tests/golden/ref_attention_*.npz are inputs/outputs of those reference kernels
    recorded on a B200 by tools/make_golden.py; tests/test_golden_cpu.py checks this
    numpy oracle against them on CPU (V bytes + params bit-identical; K bytes >= 97 %
    identical and at most one quantisation step apart -- the GPU's ex2/sin/cos.approx
    units, which the reference's --use_fast_math build uses for RoPE, cannot be
    reproduced on a CPU, so the oracle uses the correctly rounded functions; attention
    output within 2e-3);
  * tests/test_ref_attention_gpu.py runs the CUDA kernels and the reference kernels on
    the same device buffers (KV pages bit-identical, outputs within 2e-3).
The GEMM / norm / TP restatements (oracle/linear.py) have no compiled reference to run
(the reference's quantised GEMM kernels compile to empty bodies for sm_100a and need its
tuner): they are pinned by the bit-level known-answer tests in tests/test_oracle_kat.py
and the reference's own acceptance gates only -- "parity unpinned" against a reference
build for those functions.
"""
