"""Continuous-batching engine on a device: TurboMind.step through scheduler, block allocation, host marshalling, CUDA-graph
replay per batch size, prefix caching, chunked prefill.  First device run: round 2."""
import numpy as np
import pytest

from oracle import linear as OL
from oracle.numerics import BF16, F16, rn

from test_linear_gpu import DevLinear, _assert_close

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------------------
# engine level: continuous batching through TurboMind.step, prefix caching, chunked prefill
# ---------------------------------------------------------------------------------------------------------------
def _tiny_engine(monkeypatch, *, chunk=None, prefix=False):
    from lmdeploy_b200.messages import TurbomindEngineConfig
    from lmdeploy_b200.turbomind import ModelSpec
    from lmdeploy_b200.turbomind.engine import TurboMind
    # chunk=None: one prompt token per step through the decode kernel (the engine's default is chunked prefill)
    monkeypatch.setenv("B200_CHUNKED_PREFILL", str(chunk or 1))
    spec = ModelSpec("tiny", "llama", 512, 2, 4, 2, 1024, 1024, 1e4, weight_format="awq")
    cfg = TurbomindEngineConfig(max_batch_size=8, session_len=512, cache_max_entry_count=64, quant_policy=8,
                                enable_prefix_caching=prefix)
    return TurboMind(spec, cfg, seed=0)


_PROMPTS = [[5, 9, 2, 77, 3], list(range(100, 100 + 70)), [1], list(range(100, 100 + 66)) + [4, 4]]


def _generate(engine, max_new=12):
    from lmdeploy_b200.messages import GenerationConfig
    from lmdeploy_b200.turbomind.engine import batched_generate
    seqs = batched_generate(engine, _PROMPTS, GenerationConfig(max_new_tokens=max_new, ignore_eos=True))
    return [list(s.generated) for s in seqs]


def test_engine_continuous_batching_is_deterministic(monkeypatch):
    a = _generate(_tiny_engine(monkeypatch))
    b = _generate(_tiny_engine(monkeypatch))
    assert a == b and all(len(x) == 12 for x in a)


def test_engine_cuda_graph_replay_equals_eager_launches(monkeypatch):
    monkeypatch.setenv("B200_ENGINE_GRAPHS", "0")
    eager = _generate(_tiny_engine(monkeypatch), max_new=20)
    monkeypatch.setenv("B200_ENGINE_GRAPHS", "1")
    eng = _tiny_engine(monkeypatch)
    graphed = _generate(eng, max_new=20)
    assert eng._graphs, "no decode step was replayed from a graph"
    assert eager == graphed                                   # same kernels, same order, same buffers


def test_engine_prefix_caching_does_not_change_tokens(monkeypatch):
    base = _generate(_tiny_engine(monkeypatch))
    eng = _tiny_engine(monkeypatch, prefix=True)
    first, second = _generate(eng), _generate(eng)          # the second wave re-uses the published blocks
    assert eng.get_schedule_metrics().prefix_cache_hit_rate > 0
    # cached K/V were produced in a different batch composition: equal up to fp16 ties in the argmax
    for x, y, z in zip(base, first, second):
        assert x[:4] == y[:4] == z[:4]


def test_engine_chunked_prefill_matches_token_by_token(monkeypatch):
    base = _generate(_tiny_engine(monkeypatch))
    eng = _tiny_engine(monkeypatch, chunk=32)
    got = _generate(eng)
    assert eng.scheduler.tick < 70 + 12                      # the 70-token prompt no longer costs 70 steps
    for x, y in zip(base, got):
        assert x[:4] == y[:4]


def test_engine_sampling_is_a_function_of_the_seed(monkeypatch):
    """do_sample requests through the engine: the same random_seed reproduces the same tokens (whatever the batch
    composition -- every request owns its RNG state), different seeds differ, top_k = 1 is the greedy stream, and greedy
    requests sharing the batch are untouched."""
    from lmdeploy_b200.messages import GenerationConfig
    from lmdeploy_b200.turbomind.engine import batched_generate
    greedy = GenerationConfig(max_new_tokens=16, ignore_eos=True)
    def samp(seed, **kw):
        return GenerationConfig(max_new_tokens=16, ignore_eos=True, do_sample=True, top_k=kw.pop("top_k", 40), top_p=0.95,
                                temperature=1.0, random_seed=seed, **kw)
    prompts = [_PROMPTS[0], _PROMPTS[0], _PROMPTS[0], _PROMPTS[1]]
    run = lambda cfgs: [list(s.generated) for s in batched_generate(_tiny_engine(monkeypatch), prompts, cfgs)]
    a = run([samp(1), samp(2), greedy, samp(1, top_k=1)])
    b = run([samp(2), samp(1), samp(1, repetition_penalty=1.3), greedy])
    base = run([greedy] * 4)
    assert a[0] == b[1] and a[1] == b[0]                       # seed decides, not the slot or the neighbours
    assert a[0] != a[1] and a[0] != base[0]                    # 16 tokens of a random-weight model: collisions are ~impossible
    assert a[2] == base[2] and b[3] == base[3] and a[3] == base[3]      # greedy rows, and top_k = 1, give the greedy stream
    assert all(0 <= t < 1024 for row in a + b for t in row)
