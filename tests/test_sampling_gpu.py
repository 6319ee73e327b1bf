"""GPU parity of b200_sample (csrc/sampling.cu) against the oracle (oracle/sampling.py) on identical logits: the XORWOW
sequence, the size of the filtered candidate set, and -- fed the uniform number the kernel itself drew -- the sampled token.
Bars: token ids and candidate counts EXACT, except where a filter boundary falls within float rounding of the oracle's
float64 prefix sums (the kernel accumulates in fp32 in the reference's order); such rows must still pick a token inside the
oracle's candidate set."""
import ctypes as C

import numpy as np
import pytest

from oracle import sampling as S
from oracle.numerics import F16, BF16, rn

from harness import to_dev_T

pytestmark = pytest.mark.gpu


def _run(lib, logits, dtype, *, temperature, top_k, top_p, min_p, seeds, penalty=None, history=None):
    import torch
    from lmdeploy_b200 import _lib as L
    n, V = logits.shape
    ld = (V + 127) // 128 * 128
    padded = np.zeros((n, ld), np.float32)
    padded[:, :V] = logits
    padded[:, V:] = 100.0                                   # the pad columns must never be looked at
    d_logits = to_dev_T(padded, dtype)
    st_bytes = lib.b200_sampling_rng_state_bytes()
    states = torch.zeros(n * st_bytes, dtype=torch.uint8, device="cuda")
    d_seeds = torch.tensor(np.asarray(seeds, np.uint64).view(np.int64), dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    L.check(lib.b200_sampling_init_rng(states.data_ptr(), d_seeds.data_ptr(), None, n, stream), "init_rng")
    f = lambda a, t: torch.tensor(np.asarray(a), dtype=t, device="cuda")
    d_t, d_k, d_p, d_m = f(temperature, torch.float32), f(top_k, torch.int32), f(top_p, torch.float32), f(min_p, torch.float32)
    out = torch.zeros(n, dtype=torch.int32, device="cuda")
    out_u = torch.zeros(n, dtype=torch.float32, device="cuda")
    out_kept = torch.zeros(n, dtype=torch.int32, device="cuda")
    ws = torch.zeros(lib.b200_sampling_workspace_bytes(n, V), dtype=torch.uint8, device="cuda")
    p = L.SamplingParams()
    p.logits, p.ld, p.dtype, p.count, p.vocab = d_logits.data_ptr(), ld, 0 if dtype == F16 else 1, n, V
    p.temperature, p.top_k, p.top_p, p.min_p = d_t.data_ptr(), d_k.data_ptr(), d_p.data_ptr(), d_m.data_ptr()
    keep = []
    if penalty is not None:
        hs = [torch.tensor(h, dtype=torch.int32, device="cuda") for h in history]
        d_hp = torch.tensor([h.data_ptr() for h in hs], dtype=torch.int64, device="cuda")
        d_hl = torch.tensor([len(h) for h in history], dtype=torch.int32, device="cuda")
        d_pen = f(penalty, torch.float32)
        p.repetition_penalty, p.history, p.history_len = d_pen.data_ptr(), d_hp.data_ptr(), d_hl.data_ptr()
        keep = [hs, d_hp, d_hl, d_pen]
    p.rng_states, p.out_ids, p.out_uniform, p.out_kept = states.data_ptr(), out.data_ptr(), out_u.data_ptr(), out_kept.data_ptr()
    p.workspace, p.workspace_bytes, p.stream = ws.data_ptr(), ws.numel(), stream
    draws = []
    for _ in range(3):                                      # three tokens per row: the state must advance once per call
        L.check(lib.b200_sample(C.byref(p)), "b200_sample")
        torch.cuda.synchronize()
        draws.append((out.cpu().numpy().copy(), out_u.cpu().numpy().copy(), out_kept.cpu().numpy().copy()))
    del keep
    return draws


@pytest.mark.parametrize("dtype", [F16, BF16])
def test_sampling_matches_the_oracle(lib, dtype):
    rng = np.random.default_rng(7)
    V = 3000
    cfgs = [  # temperature, top_k, top_p, min_p
        (1.0, 1, 1.0, 0.0), (0.8, 50, 1.0, 0.0), (0.8, 50, 0.9, 0.0), (1.3, 40, 0.95, 0.05), (0.7, 1024, 0.8, 0.0),
        (1.0, 0, 1.0, 0.0), (1.0, 0, 0.9, 0.0), (0.6, 0, 0.7, 0.1), (1.0, 8, 0.0, 0.0), (2.0, 5000, 0.5, 0.0),
    ]
    n = len(cfgs)
    logits = rn((rng.standard_normal((n, V)) * 3).astype(np.float32), dtype)
    logits[2, :7] = logits[2, 7]                            # a tie straddling nothing in particular; ties at a top-k boundary:
    logits[1, np.argsort(-logits[1])[48:53]] = np.sort(logits[1])[-50]
    seeds = [11, 22, 33, 44, 55, 66, 77, 88, 99, 2 ** 40 + 5]
    draws = _run(lib, logits, dtype, temperature=[c[0] for c in cfgs], top_k=[c[1] for c in cfgs], top_p=[c[2] for c in cfgs],
                 min_p=[c[3] for c in cfgs], seeds=seeds)
    gens = [S.Xorwow(s) for s in seeds]
    for ids, us, kept in draws:
        for i, (t, k, tp, mp) in enumerate(cfgs):
            u = gens[i].uniform()
            assert us[i] == u, (i, us[i], u)               # curand_init(seed, 0, 0) + curand_uniform, bit for bit
            x = S.process_logits(logits[i], t)
            cand, probs = S.candidates(x, k, tp, mp)
            want = S.draw(cand, probs, u, vocabulary_order=(k <= 0 or k > V))
            assert 0 <= ids[i] < V and ids[i] in set(cand.tolist()) | {want}, (i, ids[i])
            near_boundary = abs(len(cand) - int(kept[i])) <= 1 and int(kept[i]) != len(cand)
            if not near_boundary:
                assert int(kept[i]) == len(cand), (i, int(kept[i]), len(cand))
                cum = np.cumsum(probs.astype(np.float64)) if k > 0 and k <= V else None
                # exact unless u sits within fp32 rounding of a prefix sum
                if cum is None or np.abs(cum - float(u)).min() > 1e-5:
                    assert ids[i] == want, (i, ids[i], want, float(u))


def test_repetition_penalty_and_greedy_limit(lib):
    rng = np.random.default_rng(3)
    V = 2000
    logits = rn((rng.standard_normal((3, V)) * 2).astype(np.float32), F16)
    best = int(np.argmax(logits[0]))
    history = [[best, 5, 5, 9], [1, 2, 3], [int(np.argmax(logits[2]))]]
    draws = _run(lib, logits, F16, temperature=[1.0, 1.0, 1.0], top_k=[1, 1, 1], top_p=[1.0] * 3, min_p=[0.0] * 3,
                 seeds=[1, 2, 3], penalty=[50.0, 1.0, 1.0], history=history)
    ids = draws[0][0]
    x0 = S.process_logits(logits[0], 1.0, 50.0, history[0])
    assert ids[0] == int(np.argmax(x0)) and ids[0] != best      # the penalised favourite loses
    assert ids[1] == int(np.argmax(logits[1])) and ids[2] == int(np.argmax(logits[2]))   # top_k = 1 is greedy


def test_sampled_frequencies_follow_the_probabilities(lib):
    """Size-independent property at the real vocabulary: 4096 draws from one row, top_k 20 / top_p 0.9."""
    rng = np.random.default_rng(5)
    V, n = 92544, 128
    row = rn((rng.standard_normal(V) * 2.5).astype(np.float32), F16)
    logits = np.repeat(row[None], n, 0)
    x = S.process_logits(row, 0.9)
    cand, probs = S.candidates(x, 20, 0.9, 0.0)
    counts = {}
    total = 0
    for rep in range(11):
        draws = _run(lib, logits, F16, temperature=[0.9] * n, top_k=[20] * n, top_p=[0.9] * n, min_p=[0.0] * n,
                     seeds=list(range(1000 * rep, 1000 * rep + n)))
        for ids, _, kept in draws:
            assert (kept == len(cand)).all()
            for t in ids.tolist():
                counts[t] = counts.get(t, 0) + 1
                total += 1
    assert set(counts) <= set(cand.tolist())
    for t, q in zip(cand.tolist(), probs.tolist()):
        assert abs(counts.get(t, 0) / total - q) < 4 * np.sqrt(q * (1 - q) / total) + 2e-3, (t, q, counts.get(t, 0) / total)
