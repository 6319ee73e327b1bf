"""CPU checks of the sampling oracle (oracle/sampling.py): semantics of each filter, and the distribution of the draw."""
import numpy as np

from oracle import sampling as S


def test_xorwow_is_a_pure_function_of_the_seed_and_stays_in_the_unit_interval():
    a, b = S.Xorwow(1234), S.Xorwow(1234)
    ua = [a.uniform() for _ in range(1000)]
    ub = [b.uniform() for _ in range(1000)]
    assert ua == ub and all(0.0 < u <= 1.0 for u in ua)
    assert ua != [S.Xorwow(1235).uniform() for _ in range(1000)]
    assert 0.45 < float(np.mean(ua)) < 0.55


def test_top_k_keeps_the_k_largest_in_descending_order_ties_by_id():
    x = np.array([0.5, 2.0, 2.0, -1.0, 1.0, 2.0], np.float32)
    ids, p = S.candidates(x, top_k=4)
    assert ids.tolist() == [1, 2, 5, 4] and abs(float(p.sum()) - 1) < 1e-6 and p[0] == p[1] == p[2] > p[3]
    ids, _ = S.candidates(x, top_k=2)
    assert ids.tolist() == [1, 2]                                     # exactly k survive a tie at the k-th value


def test_top_p_closes_at_the_first_prefix_above_p_and_min_p_cuts_relative_to_the_best():
    x = np.log(np.array([0.5, 0.2, 0.15, 0.1, 0.05], np.float32))
    ids, p = S.candidates(x, top_k=0, top_p=0.8)
    assert ids.tolist() == [0, 1, 2] and np.allclose(p, np.array([0.5, 0.2, 0.15]) / 0.85, atol=1e-5)
    ids, p = S.candidates(x, top_k=0, min_p=0.25)                      # 0.5 * 0.25 = 0.125: 0.1 and 0.05 leave
    assert ids.tolist() == [0, 1, 2]
    ids, p = S.candidates(x, top_k=5, top_p=0.8, min_p=0.35)           # within the top-p set, q0 * 0.35 removes 0.15
    assert ids.tolist() == [0, 1] and np.allclose(p, np.array([0.5, 0.2]) / 0.7, atol=1e-5)
    ids, _ = S.candidates(x, top_k=0, top_p=0.0)
    assert ids.tolist() == [0]


def test_repetition_penalty_and_temperature_act_before_the_filters():
    x = np.array([2.0, -2.0, 1.0], np.float32)
    y = S.process_logits(x, temperature=0.5, repetition_penalty=2.0, history=[0, 1, 1])
    assert np.allclose(y, np.array([1.0, -4.0, 1.0]) / (0.5 + 1e-6), rtol=1e-6)


def test_the_draw_follows_the_candidate_probabilities_in_either_walk_order():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(50).astype(np.float32) * 2
    ids, p = S.candidates(x, top_k=8, top_p=0.9)
    for vocab_order in (False, True):
        counts = {}
        n = 20000
        for u in rng.random(n):
            t = S.draw(ids, p, u, vocabulary_order=vocab_order)
            counts[t] = counts.get(t, 0) + 1
        assert set(counts) <= set(ids.tolist())
        for t, q in zip(ids.tolist(), p.tolist()):
            assert abs(counts.get(t, 0) / n - q) < 0.015
