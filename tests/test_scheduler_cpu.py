"""Host logic of the continuous-batching engine (no GPU): admission, block accounting, retirement, stop criteria,
and the reference-compatible config surface."""
import pytest

from lmdeploy_b200.messages import GenerationConfig, TurbomindEngineConfig
from lmdeploy_b200.turbomind.scheduler import Scheduler, Sequence


def mk(sid, prompt_len, max_new, eos=None, stop=()):
    return Sequence(session_id=sid, prompt=list(range(1, prompt_len + 1)), max_new_tokens=max_new,
                    stop_token_ids=list(stop), ignore_eos=False, eos_id=eos)


def run(sched, sampler, max_steps=10000):
    """Drive the scheduler with a fake model: sampler(seq) -> token."""
    trace = []
    steps = 0
    while sched.has_work and steps < max_steps:
        batch = sched.schedule()
        if not batch:
            break
        assert len(batch) <= sched.max_batch
        # every running sequence owns enough blocks for the token it is about to cache
        for s in batch:
            assert len(s.blocks) * sched.block_len >= s.cached + 1
        # no block is owned twice
        owned = [b for s in sched.active for b in s.blocks]
        assert len(owned) == len(set(owned))
        assert len(owned) + sched.alloc.num_free == sched.alloc.total
        trace.append([s.session_id for s in batch])
        sched.update(batch, [sampler(s) for s in batch])
        steps += 1
    sched.schedule()          # final retirement
    return trace


def test_prompt_is_fed_one_token_per_step_then_generation():
    sched = Scheduler(max_batch=4, block_len=64, num_blocks=8, session_len=1024)
    s = mk(1, prompt_len=5, max_new=3)
    sched.add(s)
    fed = []
    run(sched, lambda q: (fed.append(q.next_input()), 100 + len(q.generated))[1])
    assert fed[:5] == [1, 2, 3, 4, 5]            # the prompt, in order
    assert fed[5:] == [100, 101]                 # then its own samples
    assert s.generated == [100, 101, 102] and s.finish_reason == "length"
    assert s.cached == 5 + 3 - 1 + 0 or s.cached == 7
    assert sched.alloc.num_free == 8             # all blocks returned


def test_continuous_batching_admits_when_slots_free_up():
    sched = Scheduler(max_batch=2, block_len=64, num_blocks=16, session_len=1024)
    for i, n in enumerate([3, 6, 2, 4]):
        sched.add(mk(i, prompt_len=2, max_new=n))
    trace = run(sched, lambda q: 7)
    assert max(len(t) for t in trace) == 2
    first_seen = {sid: next(i for i, t in enumerate(trace) if sid in t) for sid in range(4)}
    assert first_seen[0] == 0 and first_seen[1] == 0 and first_seen[2] > 0 and first_seen[3] > first_seen[2] - 1
    assert sched.alloc.num_free == 16 and not sched.sessions


def test_block_growth_and_exhaustion():
    sched = Scheduler(max_batch=4, block_len=64, num_blocks=3, session_len=4096)
    a, b = mk(1, 60, 200), mk(2, 60, 200)
    sched.add(a), sched.add(b)
    run(sched, lambda q: 9)
    # 3 blocks cannot hold 2 x 260 tokens: somebody is stopped by KV exhaustion, nobody deadlocks
    assert a.finished and b.finished
    assert "length" in (a.finish_reason, b.finish_reason)
    assert sched.alloc.num_free == 3


def test_stop_criteria_eos_and_stop_ids_and_cancel():
    sched = Scheduler(max_batch=4, block_len=64, num_blocks=8, session_len=1024)
    a, b, c = mk(1, 2, 50, eos=2), mk(2, 2, 50, stop=[5]), mk(3, 2, 50)
    for s in (a, b, c):
        sched.add(s)
    toks = {1: iter([9, 9, 2, 9]), 2: iter([9, 5, 9]), 3: iter([1] * 100)}
    steps = [0]

    def sampler(q):
        steps[0] += 1
        if steps[0] == 12:
            sched.cancel(3)
        return next(toks[q.session_id]) if q.fed >= len(q.prompt) - 1 else 0
    run(sched, sampler)
    assert a.finish_reason == "stop" and a.generated[-1] == 2
    assert b.finish_reason == "stop" and b.generated[-1] == 5
    assert c.finish_reason == "cancel"


def test_session_len_clamps_and_duplicate_session_rejected():
    sched = Scheduler(max_batch=2, block_len=64, num_blocks=64, session_len=100)
    s = mk(1, 90, 500)
    sched.add(s)
    assert s.max_new_tokens == 10
    with pytest.raises(KeyError):
        sched.add(mk(1, 3, 3))
    with pytest.raises(ValueError):
        sched.add(Sequence(2, [], 1, [], False, None))


def test_engine_config_validation_matches_reference_rules():
    # lmdeploy/messages.py:333-359 __post_init__ rules that touch this path
    TurbomindEngineConfig(quant_policy=8, tp=2, cache_block_seq_len=128)
    with pytest.raises(ValueError):
        TurbomindEngineConfig(quant_policy=3)
    with pytest.raises(ValueError):
        TurbomindEngineConfig(tp=3)
    with pytest.raises(ValueError):
        TurbomindEngineConfig(cache_max_entry_count=0)
    with pytest.raises(ValueError):
        TurbomindEngineConfig(dtype="float32")
    with pytest.raises(NotImplementedError):
        GenerationConfig(do_sample=True, top_k=40)
    GenerationConfig(max_new_tokens=8, top_k=1, ignore_eos=True)


def test_product_fails_loudly_without_library(monkeypatch, tmp_path):
    from lmdeploy_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.B200LibraryError):
        _lib.load()
