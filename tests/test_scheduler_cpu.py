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
        steps += 1
        if not batch:
            continue          # e.g. a tick that only retired / preempted; the engine loop simply calls step() again
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
    assert steps < max_steps, "the scheduler stopped making progress"
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
    GenerationConfig(do_sample=True, top_k=40)                 # device sampling stage (csrc/sampling.cu)
    with pytest.raises(ValueError):
        GenerationConfig(do_sample=True, top_k=2000)           # the top-k kernel's limit, as in the reference
    GenerationConfig(max_new_tokens=8, top_k=1, ignore_eos=True)


def test_product_fails_loudly_without_library(monkeypatch, tmp_path):
    from lmdeploy_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.B200LibraryError):
        _lib.load()


# ---------------------------------------------------------------------------------------------------------------
# prefix caching
# ---------------------------------------------------------------------------------------------------------------
class FakeKvModel:
    """Stands in for the decode step: `store[block][offset]` plays the KV page (it records the token whose K/V would be
    written there) and the sampled token is a hash of everything the sequence can see through ITS block table -- so a
    wrong, stale or shared-but-overwritten block changes the output."""

    def __init__(self, num_blocks, block_len, vocab=50000):
        self.store = [[None] * block_len for _ in range(num_blocks)]
        self.block_len, self.vocab = block_len, vocab
        self.fed = 0

    def step(self, batch):
        out = []
        for s in batch:
            pos = s.cached
            self.store[s.blocks[pos // self.block_len]][pos % self.block_len] = s.next_input()
            self.fed += 1
        for s in batch:
            h = 1469598103934665603
            for p in range(s.cached + 1):
                t = self.store[s.blocks[p // self.block_len]][p % self.block_len]
                assert t is not None, "read of a KV slot nobody wrote"
                h = ((h ^ (t + 1)) * 1099511628211) % (1 << 64)
            out.append(h % self.vocab)
        return out


def generate(prompts, *, prefix_caching, num_blocks=64, block_len=4, max_batch=4, max_new=6, waves=1):
    sched = Scheduler(max_batch=max_batch, block_len=block_len, num_blocks=num_blocks, session_len=4096,
                      prefix_caching=prefix_caching)
    model = FakeKvModel(num_blocks, block_len)
    results, sid = [], 0
    for _ in range(waves):
        seqs = []
        for p in prompts:
            sid += 1
            seqs.append(Sequence(session_id=sid, prompt=list(p), max_new_tokens=max_new, stop_token_ids=[],
                                 ignore_eos=True, eos_id=None))
            sched.add(seqs[-1])
        while sched.has_work:
            batch = sched.schedule()
            assert batch
            # a block may be shared, but then only below every holder's write position
            for s in batch:
                wb = s.blocks[s.cached // block_len]
                assert sum(wb in o.blocks for o in sched.active) == 1, "a shared block is about to be written"
            sched.update(batch, model.step(batch))
        sched.schedule()
        results.append([list(s.generated) for s in seqs])
    return results, model, sched


def test_prefix_cache_reuses_blocks_and_preserves_outputs():
    system = list(range(100, 100 + 11))                       # 11 shared tokens: 2 full blocks of 4 + 3
    prompts = [system + [7, 8, 9], system + [7, 1], system[:9] + [5, 5, 5, 5], [3, 1, 4, 1, 5]]
    base, m0, _ = generate(prompts, prefix_caching=False, waves=2)
    got, m1, sched = generate(prompts, prefix_caching=True, waves=2)
    assert got == base                                         # identical generations, both waves
    assert m1.fed < m0.fed                                     # and fewer tokens fed
    # second wave: every prompt finds all of its full blocks except the one holding its last token
    expect_reuse = sum(((len(p) - 1) // 4) * 4 for p in prompts)
    assert m0.fed - m1.fed >= expect_reuse
    met = sched.metrics()
    assert 0 < met["prefix_cache_hit_rate"] < 1 and met["active_seqs"] == 0 and met["active_blocks"] == 0
    assert met["free_blocks"] + met["cached_blocks"] == met["total_blocks"]


def test_prefix_cache_never_covers_the_last_prompt_token():
    prompt = list(range(1, 9))                                 # exactly 2 blocks of 4
    sched = Scheduler(max_batch=2, block_len=4, num_blocks=16, session_len=64, prefix_caching=True)
    model = FakeKvModel(16, 4)
    outs = []
    for sid in (1, 2):
        s = Sequence(session_id=sid, prompt=list(prompt), max_new_tokens=2, stop_token_ids=[], ignore_eos=True, eos_id=None)
        sched.add(s)
        while sched.has_work:
            b = sched.schedule()
            sched.update(b, model.step(b))
        outs.append((s.generated, s.reused_tokens))
    assert outs[0][0] == outs[1][0]
    assert outs[0][1] == 0 and outs[1][1] == 4                 # block 0 only: block 1 holds the last prompt token


def test_prefix_cache_eviction_under_pressure_keeps_accounting_exact():
    # pool of 6 blocks, block_len 4: cached blocks of finished requests must be reclaimed for new, unrelated ones
    sched = Scheduler(max_batch=1, block_len=4, num_blocks=6, session_len=64, prefix_caching=True)
    model = FakeKvModel(6, 4)
    gens = {}
    for sid in range(1, 8):
        prompt = [sid * 10 + i for i in range(9)]              # distinct prompts, 3 blocks each
        s = Sequence(session_id=sid, prompt=prompt, max_new_tokens=3, stop_token_ids=[], ignore_eos=True, eos_id=None)
        sched.add(s)
        while sched.has_work:
            b = sched.schedule()
            assert b, "admission starved although idle cached blocks could be evicted"
            sched.update(b, model.step(b))
        sched.schedule()
        assert s.finish_reason == "length" and len(s.generated) == 3
        gens[sid] = list(s.generated)
        m = sched.metrics()
        assert m["free_blocks"] + m["cached_blocks"] == 6 and m["active_blocks"] == 0
    # the same prompts without caching give the same tokens
    base, _, _ = generate([[sid * 10 + i for i in range(9)] for sid in range(1, 8)], prefix_caching=False, num_blocks=6,
                          max_batch=1, max_new=3)
    assert [gens[i] for i in range(1, 8)] == base[0]


def test_concurrent_identical_prompts_do_not_share_a_block_being_written():
    prompts = [[9, 8, 7, 6, 5, 4, 3, 2, 1]] * 3
    base, _, _ = generate(prompts, prefix_caching=False, max_batch=3)
    got, _, sched = generate(prompts, prefix_caching=True, max_batch=3)
    assert got == base
    # all three ran concurrently from scratch (nothing was published when they were admitted): no reuse
    assert sched.prefix.hits == 0


def test_cancel_and_retire_return_references():
    sched = Scheduler(max_batch=2, block_len=4, num_blocks=8, session_len=64, prefix_caching=True)
    model = FakeKvModel(8, 4)
    a = Sequence(session_id=1, prompt=list(range(1, 10)), max_new_tokens=50, stop_token_ids=[], ignore_eos=True, eos_id=None)
    sched.add(a)
    for _ in range(9):
        b = sched.schedule()
        sched.update(b, model.step(b))
    c = Sequence(session_id=2, prompt=list(range(1, 10)), max_new_tokens=50, stop_token_ids=[], ignore_eos=True, eos_id=None)
    sched.add(c)
    b = sched.schedule()
    assert c.reused_tokens == 8 and c.blocks[:2] == a.blocks[:2]
    sched.update(b, model.step(b))
    sched.cancel(1)
    b = sched.schedule()                                       # retires a; c keeps the shared blocks alive
    assert [s.session_id for s in b] == [2]
    assert all(sched.prefix.refs[x] == 1 for x in c.blocks[:2])
    sched.cancel(2)
    sched.schedule()
    m = sched.metrics()
    assert m["active_blocks"] == 0 and m["free_blocks"] + m["cached_blocks"] == 8 and m["total_seqs"] == 0


# ---------------------------------------------------------------------------------------------------------------
# chunked prefill (several prompt tokens of a sequence per step, mixed with decoding sequences)
# ---------------------------------------------------------------------------------------------------------------
class FakeChunkModel(FakeKvModel):
    """FakeKvModel for steps in which a sequence feeds `step_tokens` tokens: writes them at positions cached ..
    cached + n - 1 and samples from what the LAST of them can see (causal)."""

    def step(self, batch):
        rows = 0
        for s in batch:
            toks = s.next_inputs()
            assert len(toks) == s.step_tokens >= 1
            for i, t in enumerate(toks):
                pos = s.cached + i
                self.store[s.blocks[pos // self.block_len]][pos % self.block_len] = t
            rows += len(toks)
            self.fed += len(toks)
        self.max_rows = max(getattr(self, "max_rows", 0), rows)
        out = []
        for s in batch:
            h = 1469598103934665603
            for p in range(s.cached + s.step_tokens):
                t = self.store[s.blocks[p // self.block_len]][p % self.block_len]
                assert t is not None
                h = ((h ^ (t + 1)) * 1099511628211) % (1 << 64)
            out.append(h % self.vocab)
        return out


def run_chunked(prompts, *, prefill_chunk, token_budget=None, prefix_caching=False, max_batch=4, block_len=4, num_blocks=64,
                max_new=5):
    sched = Scheduler(max_batch=max_batch, block_len=block_len, num_blocks=num_blocks, session_len=4096,
                      prefix_caching=prefix_caching, prefill_chunk=prefill_chunk, token_budget=token_budget)
    model = FakeChunkModel(num_blocks, block_len)
    seqs = [Sequence(session_id=i + 1, prompt=list(p), max_new_tokens=max_new, stop_token_ids=[], ignore_eos=True,
                     eos_id=None) for i, p in enumerate(prompts)]
    for s in seqs:
        sched.add(s)
    steps = 0
    while sched.has_work:
        batch = sched.schedule()
        assert batch
        assert sum(s.step_tokens for s in batch) <= sched.token_budget
        for s in batch:
            assert len(s.blocks) * block_len >= s.cached + s.step_tokens
            assert s.step_tokens == 1 or s.fed + s.step_tokens <= len(s.prompt)
        sched.update(batch, model.step(batch))
        steps += 1
    sched.schedule()
    assert sched.alloc.num_free + (sched.prefix.num_idle if sched.prefix else 0) == num_blocks
    return [list(s.generated) for s in seqs], steps, model


PROMPTS = [list(range(10, 10 + 23)), [7], list(range(50, 50 + 9)), list(range(10, 10 + 17)) + [99, 98]]


def test_chunked_prefill_gives_the_same_tokens_in_fewer_steps():
    base, steps1, _ = run_chunked(PROMPTS, prefill_chunk=1)
    for chunk in (2, 5, 8, 64):
        got, steps, model = run_chunked(PROMPTS, prefill_chunk=chunk, token_budget=16)
        assert got == base, chunk
        assert steps < steps1
        assert model.max_rows <= 16
    # budget == max_batch leaves no room for longer chunks once the batch is full, but still terminates correctly
    got, _, model = run_chunked(PROMPTS, prefill_chunk=8, token_budget=4)
    assert got == base and model.max_rows <= 4


def test_chunked_prefill_with_prefix_cache_publishes_every_filled_block():
    base, _, _ = run_chunked(PROMPTS, prefill_chunk=1)
    got, _, _ = run_chunked(PROMPTS, prefill_chunk=8, token_budget=32, prefix_caching=True)
    assert got == base
    # second identical wave on one scheduler: the shared 16-token prefix of prompts 0 and 3 and every full block of each
    # prompt must be found, although they were written several blocks per step
    sched = Scheduler(max_batch=2, block_len=4, num_blocks=64, session_len=4096, prefix_caching=True, prefill_chunk=8,
                      token_budget=16)
    model = FakeChunkModel(64, 4)
    outs = []
    for wave in range(2):
        seqs = [Sequence(session_id=wave * 10 + i, prompt=list(p), max_new_tokens=3, stop_token_ids=[], ignore_eos=True,
                         eos_id=None) for i, p in enumerate(PROMPTS)]
        for s in seqs:
            sched.add(s)
        while sched.has_work:
            b = sched.schedule()
            sched.update(b, model.step(b))
        sched.schedule()
        outs.append([list(s.generated) for s in seqs])
        reused = [s.reused_tokens for s in seqs]
    assert outs[0] == outs[1]
    assert reused == [((len(p) - 1) // 4) * 4 for p in PROMPTS]


def test_chunk_never_swallows_the_last_prompt_token_early():
    # the step that feeds the last prompt token is the one whose sample is kept; a chunk may end exactly there
    sched = Scheduler(max_batch=1, block_len=4, num_blocks=8, session_len=64, prefill_chunk=4, token_budget=8)
    s = Sequence(session_id=1, prompt=[1, 2, 3, 4, 5, 6], max_new_tokens=2, stop_token_ids=[], ignore_eos=True, eos_id=None)
    sched.add(s)
    sizes = []
    model = FakeChunkModel(8, 4)
    while sched.has_work:
        b = sched.schedule()
        sizes.append(b[0].step_tokens)
        sched.update(b, model.step(b))
    assert sizes == [4, 2, 1] and len(s.generated) == 2


def test_marshal_step_describes_the_batch_the_way_the_c_abi_reads_it():
    """Re-derive every sequence's (new tokens, history, block table) from the marshalled arrays with the C ABI's own
    conventions (b200_attention_params: rows [cu_q_len[b], cu_q_len[b+1]), k_len = cu_k_len[b+1] - cu_k_len[b]
    INCLUDING the new tokens, blocks [cu_block_nums[b], cu_block_nums[b+1]))."""
    from lmdeploy_b200.turbomind.engine import marshal_step
    sched = Scheduler(max_batch=4, block_len=4, num_blocks=64, session_len=4096, prefill_chunk=6, token_budget=12)
    model = FakeChunkModel(64, 4)
    for i, p in enumerate(PROMPTS):
        sched.add(Sequence(session_id=i + 1, prompt=list(p), max_new_tokens=4, stop_token_ids=[], ignore_eos=True, eos_id=None))
    seen_multi = False
    while sched.has_work:
        batch = sched.schedule()
        st = marshal_step(batch, lambda b: 1000 + 16 * b)
        assert st["token_num"] == len(st["token_ids"]) == st["cu_q_len"][-1] <= 12
        for i, s in enumerate(batch):
            rows = st["token_ids"][st["cu_q_len"][i]:st["cu_q_len"][i + 1]]
            k_len = st["cu_k_len"][i + 1] - st["cu_k_len"][i]
            blocks = st["block_ptrs"][st["cu_block_nums"][i]:st["cu_block_nums"][i + 1]]
            assert rows == s.next_inputs() and len(rows) == s.step_tokens
            assert k_len - len(rows) == s.cached                       # history_len of the attention call
            assert blocks == [1000 + 16 * b for b in s.blocks]
            assert len(blocks) * 4 >= k_len                            # every key position has a page
            seen_multi |= len(rows) > 1
        assert st["max_k_len"] == max(st["cu_k_len"][i + 1] - st["cu_k_len"][i] for i in range(len(batch)))
        sched.update(batch, model.step(batch))
    assert seen_multi


from fake_tokenizer import FakeTokenizer as _FakeTokenizer  # noqa: E402


def test_stop_words_become_stop_token_ids_with_the_reference_rule():
    """Against tests/golden/ref_stop_words.json: the reference's OWN GenerationConfig.convert_stop_bad_words_to_ids /
    update_from_hf_gen_cfg and Tokenizer.indexes_containing_token, executed by tools/make_stopwords_golden.py over the
    toy vocabulary of tests/fake_tokenizer.py (substring matches, the ' ' -> sentencepiece marker rule, narrowing of
    more than five candidates, multi-token phrases dropped, ids beyond the vocabulary)."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_stop_words.json")))
    assert g["vocab"] == _FakeTokenizer.VOCAB
    tok = _FakeTokenizer()
    assert len(g["stop_words"]) >= 8
    for case in g["stop_words"]:
        cfg = GenerationConfig(stop_words=case.get("stop_words"), stop_token_ids=case.get("stop_token_ids"))
        cfg.convert_stop_bad_words_to_ids(tok)
        assert cfg.stop_token_ids == case["result"], case
    for case in g["eos"]:
        cfg = GenerationConfig(stop_token_ids=case["stop_token_ids"])
        cfg.update_from_hf_gen_cfg(case["generation_config"], case["tokenizer_eos"])
        assert cfg.stop_token_ids == case["result"], case
    with pytest.raises(NotImplementedError):                  # the one deliberate refusal: bad words need a logits processor
        GenerationConfig(bad_words=["x"]).convert_stop_bad_words_to_ids(tok)


def test_pipeline_prompt_forms_without_a_device():
    """Pipeline._encode / single-vs-batch detection: token ids, text, chat messages (no engine is constructed)."""
    from lmdeploy_b200.api import Pipeline

    class Tok(_FakeTokenizer):
        def apply_chat_template(self, messages, add_generation_prompt=True, tokenize=True):
            assert add_generation_prompt and tokenize
            out = [0]
            for m in messages:
                out += self.encode(m["content"])
            return out + [1]

    p = Pipeline.__new__(Pipeline)
    p.tokenizer = Tok()
    assert p._encode([5, 6, 7]) == [5, 6, 7]
    assert p._encode("!!") == [5]
    assert p._encode([{"role": "user", "content": "!"}, {"role": "assistant", "content": "!!"}]) == [0, 4, 5, 1]
    p.tokenizer = None
    assert p._encode((1, 2)) == [1, 2]
    with pytest.raises(ValueError):
        p._encode("text")
    with pytest.raises(ValueError):
        p._encode([{"role": "user", "content": "x"}])


# ---------------------------------------------------------------------------------------------------------------
# randomized: prefix caching x chunked prefill x memory pressure x cancellation
# ---------------------------------------------------------------------------------------------------------------
def _drive(prompts, max_new, *, num_blocks, max_batch, prefill_chunk, token_budget, prefix_caching, cancel_at, arrivals,
           block_len=4):
    """Run a workload to completion; returns {session: (generated, finish_reason)}.  Sequences arrive at given steps,
    one session may be cancelled at a given step.  Invariants are asserted every step."""
    sched = Scheduler(max_batch=max_batch, block_len=block_len, num_blocks=num_blocks, session_len=4096,
                      prefix_caching=prefix_caching, prefill_chunk=prefill_chunk, token_budget=token_budget)
    model = FakeChunkModel(num_blocks, block_len)
    seqs = {}
    step = 0
    pending = sorted(range(len(prompts)), key=lambda i: arrivals[i])
    while pending or sched.has_work:
        while pending and arrivals[pending[0]] <= step:
            i = pending.pop(0)
            seqs[i] = Sequence(session_id=i, prompt=list(prompts[i]), max_new_tokens=max_new[i], stop_token_ids=[],
                               ignore_eos=True, eos_id=None)
            sched.add(seqs[i])
        if cancel_at is not None and cancel_at[1] == step:
            sched.cancel(cancel_at[0])
        batch = sched.schedule()
        if not batch:
            if not pending and not sched.has_work:
                break
            step += 1
            assert step < 5000, "scheduler makes no progress"
            continue
        assert len(batch) <= max_batch and sum(s.step_tokens for s in batch) <= sched.token_budget
        owners = {}
        for s in sched.active:
            for j, b in enumerate(s.blocks):
                owners.setdefault(b, []).append((s, j))
        for b, who in owners.items():
            if len(who) > 1:        # shared: must be a published block, entirely below every holder's write position
                assert sched.prefix is not None and b in sched.prefix.node_of
                for s, j in who:
                    assert (j + 1) * block_len <= s.cached
        for s in batch:
            assert len(s.blocks) * block_len >= s.cached + s.step_tokens
        idle = sched.prefix.num_idle if sched.prefix else 0
        assert len(owners) + sched.alloc.num_free + idle == num_blocks
        sched.update(batch, model.step(batch))
        step += 1
        assert step < 5000
    sched.schedule()
    idle = sched.prefix.num_idle if sched.prefix else 0
    assert sched.alloc.num_free + idle == num_blocks and not sched.sessions
    return {i: (list(s.generated), s.finish_reason) for i, s in seqs.items()}


def test_randomized_workloads_agree_with_the_plain_scheduler():
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    prompt = st.lists(st.integers(0, 5), min_size=1, max_size=22)        # tiny alphabet: many shared prefixes

    @settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck))
    @given(prompts=st.lists(prompt, min_size=1, max_size=6), data=st.data())
    def run(prompts, data):
        n = len(prompts)
        max_new = [data.draw(st.integers(0, 5)) for _ in range(n)]
        arrivals = [data.draw(st.integers(0, 12)) for _ in range(n)]
        max_batch = data.draw(st.integers(1, 4))
        chunk = data.draw(st.sampled_from([1, 2, 5, 64]))
        budget = data.draw(st.integers(max_batch, 12))
        roomy = sum((len(p) + m + 3) // 4 + 1 for p, m in zip(prompts, max_new)) + 2
        base = _drive(prompts, max_new, num_blocks=roomy, max_batch=max_batch, prefill_chunk=1, token_budget=None,
                      prefix_caching=False, cancel_at=None, arrivals=arrivals)
        got = _drive(prompts, max_new, num_blocks=roomy, max_batch=max_batch, prefill_chunk=chunk, token_budget=budget,
                     prefix_caching=True, cancel_at=None, arrivals=arrivals)
        assert got == base
        # under memory pressure and with a cancellation the run must still terminate with exact accounting, and every
        # sequence that finished normally must have produced the same tokens
        tight = max((len(p) + m + 3) // 4 + 1 for p, m in zip(prompts, max_new)) + data.draw(st.integers(0, 3))
        cancel = (data.draw(st.integers(0, n - 1)), data.draw(st.integers(0, 20))) if data.draw(st.booleans()) else None
        pressed = _drive(prompts, max_new, num_blocks=tight, max_batch=max_batch, prefill_chunk=chunk, token_budget=budget,
                         prefix_caching=True, cancel_at=cancel, arrivals=arrivals)
        for i, (gen, why) in pressed.items():
            if cancel is not None and i == cancel[0] and why == "cancel":
                assert gen == base[i][0][:len(gen)]          # cut short by the cancellation: a prefix
            else:
                # every sequence fits the pool on its own, so memory pressure only preempts (the victim refills its
                # cache later): nobody is cut short, everybody produces exactly the unpressed tokens
                assert (gen, why) == base[i], (i, gen, why, base[i])

    run()


def test_config_keyword_surface_matches_the_reference():
    """Every field of the reference's TurbomindEngineConfig / GenerationConfig (tests/golden/ref_config_fields.json, read
    from lmdeploy/messages.py by tools/make_config_golden.py) is accepted here with the same default, so user code that
    constructs them switches packages unchanged; the one deliberate difference is named."""
    import dataclasses
    import json
    import os
    import lmdeploy_b200.messages as M
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_config_fields.json")))
    deliberate = {("TurbomindEngineConfig", "communicator"): "'native'"}     # the NVLink peer-memory kernels, not NCCL
    for cls, fields in ref.items():
        ours = {f.name: f for f in dataclasses.fields(getattr(M, cls))}
        assert set(fields) <= set(ours), sorted(set(fields) - set(ours))
        for name, default in fields.items():
            got = repr(ours[name].default)
            want = deliberate.get((cls, name), default)
            assert got.replace('"', "'") == str(want).replace('"', "'"), (cls, name, got, want)
    M.TurbomindEngineConfig(communicator="nccl")            # accepted: selects nothing else
    for kw in (dict(cp=2), dict(ep=2), dict(nnodes=2), dict(attn_tp_size=2), dict(attn_dp_size=2)):
        with pytest.raises(NotImplementedError):
            M.TurbomindEngineConfig(**kw)
    for kw in (dict(response_format={"type": "json_schema"}), dict(logits_processors=[print]), dict(return_ppl=True),
               dict(output_logits="all"), dict(output_last_hidden_state="all"),
               dict(repetition_ngram_size=3, repetition_ngram_threshold=2)):
        with pytest.raises(NotImplementedError):
            M.GenerationConfig(**kw)
    # the reference's own validation (messages.py:195-205), raised as an error that is both ValueError and AssertionError
    for kw in (dict(n=0), dict(top_p=1.5), dict(top_k=-1), dict(temperature=2.5), dict(min_p=-0.1)):
        with pytest.raises(AssertionError):
            M.GenerationConfig(**kw)
        with pytest.raises(ValueError):
            M.GenerationConfig(**kw)
    g = M.GenerationConfig(repetition_ngram_size=3)               # threshold 0: the pair is reset, as the reference does
    assert (g.repetition_ngram_size, g.repetition_ngram_threshold) == (0, 0)
    with pytest.raises(AssertionError):
        M.TurbomindEngineConfig(cache_max_entry_count=0)


class _FakeEngine:
    """The three things the pipeline layer uses of TurboMind -- scheduler, eos_id, step(), new_session_id() -- over the
    fake KV model, so Pipeline.infer / stream_infer run on CPU."""

    def __init__(self, **sched_kw):
        self.scheduler = Scheduler(max_batch=3, block_len=4, num_blocks=64, session_len=256, **sched_kw)
        self.model = FakeChunkModel(64, 4, vocab=12)             # tiny vocabulary: stop tokens get hit
        self.eos_id = None
        self._sid = 0

    def new_session_id(self):
        self._sid += 1
        return self._sid

    def step(self):
        batch = self.scheduler.schedule()
        if batch:
            return self.scheduler.update(batch, self.model.step(batch))
        return []


def _fake_pipeline(**sched_kw):
    from lmdeploy_b200.api import Pipeline
    p = Pipeline.__new__(Pipeline)
    p.engine, p.tokenizer = _FakeEngine(**sched_kw), None
    return p


def test_pipeline_infer_and_stream_infer_agree():
    prompts = [[3, 1, 4, 1, 5, 9, 2, 6], [2, 7], [1, 8, 2, 8, 1, 8]]
    cfgs = [GenerationConfig(max_new_tokens=n, ignore_eos=True) for n in (5, 9, 1)]
    full = _fake_pipeline().infer(prompts, cfgs)
    assert [r.index for r in full] == [0, 1, 2] and [r.generate_token_len for r in full] == [5, 9, 1]
    assert all(r.finish_reason == "length" and r.input_token_len == len(p) for r, p in zip(full, prompts))
    single = _fake_pipeline().infer(prompts[0], cfgs[0])
    assert single.token_ids == full[0].token_ids                # a single prompt returns a single Response
    # streamed: the concatenation of a prompt's chunks is its full answer; exactly one final chunk per prompt
    chunks = list(_fake_pipeline().stream_infer(prompts, gen_config=cfgs))
    for i in range(3):
        mine = [c for c in chunks if c.index == i]
        assert sum((c.token_ids for c in mine), []) == full[i].token_ids
        assert [c.finish_reason for c in mine][:-1] == [None] * (len(mine) - 1) and mine[-1].finish_reason == "length"
        assert all(c.token_ids for c in mine[:-1]) and mine[-1].token_ids == [] and mine[-1].text == ""      # closing chunk
        assert [c.generate_token_len for c in mine[:-1]] == list(range(1, len(mine))) and \
            mine[-1].generate_token_len == full[i].generate_token_len
    # non-streaming mode of stream_infer: one content chunk with everything + the closing chunk, per prompt
    finals = list(_fake_pipeline().stream_infer(prompts, gen_config=cfgs, stream_response=False))
    assert sorted((r.index, tuple(r.token_ids), r.finish_reason) for r in finals if r.finish_reason is None) == \
        [(i, tuple(full[i].token_ids), None) for i in range(3)]
    assert sorted((r.index, r.finish_reason) for r in finals if r.finish_reason) == [(i, "length") for i in range(3)]
    # per-prompt stop tokens; a cancelled request reports 'abort'
    stop = full[1].token_ids[3]
    # the stop token ends the sequence and stays out of the response (async_engine.py:717-720), unless asked for
    cut = full[1].token_ids.index(stop)
    r = _fake_pipeline().infer([prompts[1]], [GenerationConfig(max_new_tokens=9, stop_token_ids=[stop])])[0]
    assert r.finish_reason == "stop" and r.token_ids == full[1].token_ids[:cut] and r.generate_token_len == cut
    r = _fake_pipeline().infer([prompts[1]], [GenerationConfig(max_new_tokens=9, stop_token_ids=[stop],
                                                               include_stop_str_in_output=True)])[0]
    assert r.finish_reason == "stop" and r.token_ids == full[1].token_ids[:cut + 1] and r.generate_token_len == cut + 1
    ch = list(_fake_pipeline().stream_infer([prompts[1]], gen_config=[GenerationConfig(max_new_tokens=9, stop_token_ids=[stop])]))
    assert sum((c.token_ids for c in ch), []) == full[1].token_ids[:cut] and ch[-1].finish_reason == "stop" and ch[-1].token_ids == []
    ch = list(_fake_pipeline().stream_infer([prompts[1]], gen_config=[GenerationConfig(max_new_tokens=9, stop_token_ids=[stop],
                                                                                    include_stop_str_in_output=True)]))
    assert ch[-1].token_ids == [stop] and ch[-1].generate_token_len == cut + 1
    # ignore_eos switches every stop id off (turbomind.py _get_generation_config hands stop_ids over only without it)
    r = _fake_pipeline().infer([prompts[1]], [GenerationConfig(max_new_tokens=9, stop_token_ids=[stop], ignore_eos=True)])[0]
    assert r.finish_reason == "length" and r.token_ids == full[1].token_ids
    with pytest.raises(ValueError):
        _fake_pipeline().infer(prompts, cfgs[:2])
    # prefix cache: the second identical request reports its reused tokens
    p = _fake_pipeline(prefix_caching=True)
    a = p.infer([prompts[0]], cfgs[0])[0]
    b = p.infer([prompts[0]], cfgs[0])[0]
    assert a.token_ids == b.token_ids and a.cached_tokens == 0 and b.cached_tokens == 4


def test_options_that_would_change_greedy_tokens_are_refused():
    GenerationConfig(temperature=0.3, top_p=0.8, top_k=40, min_p=0.1)          # sampling knobs are inert without do_sample
    GenerationConfig(do_sample=True, top_k=1)                                   # top_k = 1 is greedy
    GenerationConfig(repetition_penalty=1.1)        # reset to 1.0 for greedy requests by the reference (_determine_gen_config)
    GenerationConfig(do_sample=True, top_k=40, repetition_penalty=1.1)         # sampled on the device
    for kw in (dict(min_new_tokens=4),):
        with pytest.raises(NotImplementedError):
            GenerationConfig(**kw)


def test_unbounded_generation_stops_at_the_session_length():
    p = _fake_pipeline()
    p.engine.scheduler.session_len = 20
    r = p.infer([[1, 2, 3, 4, 5]], GenerationConfig(max_new_tokens=None, ignore_eos=True))[0]
    assert r.finish_reason == "length" and r.input_token_len + r.generate_token_len == 20


def test_a_prompt_that_fills_the_session_is_rejected_alone():
    p = _fake_pipeline()
    p.engine.scheduler.session_len = 12
    ok, bad = [1, 2, 3], list(range(12))
    res = p.infer([ok, bad], GenerationConfig(max_new_tokens=3, ignore_eos=True))
    assert res[0].finish_reason == "length" and len(res[0].token_ids) == 3
    assert res[1].finish_reason == "error" and res[1].error_code == "input_length" and "session_len 12" in res[1].error_message
    assert res[1].token_ids == [] and p.engine.scheduler.metrics()["total_seqs"] == 0
    # instance level: the reference's status code
    from lmdeploy_b200.messages import ResponseType
    from lmdeploy_b200.turbomind.engine import TurboMindInstance
    outs = list(TurboMindInstance(p.engine).stream_infer(99, bad, GenerationConfig(max_new_tokens=3)))
    assert [o.status for o in outs] == [ResponseType.INPUT_LENGTH_ERROR]


def test_streamed_text_and_engine_outputs_are_incremental():
    from lmdeploy_b200.messages import ResponseType
    from lmdeploy_b200.turbomind.engine import TurboMindInstance
    prompt, cfg = [2, 7, 1], GenerationConfig(max_new_tokens=8, ignore_eos=True, skip_special_tokens=False)
    p = _fake_pipeline()
    p.tokenizer = _FakeTokenizer()
    whole = p.infer([prompt], cfg)[0]
    assert whole.text == p.tokenizer.decode(whole.token_ids) and len(whole.token_ids) == 8
    p2 = _fake_pipeline()
    p2.tokenizer = _FakeTokenizer()
    chunks = list(p2.stream_infer([prompt], gen_config=cfg))
    assert "".join(c.text for c in chunks) == whole.text and sum((c.token_ids for c in chunks), []) == whole.token_ids
    # engine level: EngineOutput.token_ids are the NEW ids of the iteration (lmdeploy/messages.py:721-743), num_token the total
    outs = list(TurboMindInstance(_fake_pipeline().engine).stream_infer(7, prompt, cfg))
    assert [o.status for o in outs[:-1]] == [ResponseType.SUCCESS] * (len(outs) - 1) and outs[-1].status == ResponseType.FINISH
    assert sum((o.token_ids for o in outs), []) == whole.token_ids
    assert [o.num_token for o in outs] == sorted(o.num_token for o in outs) and outs[-1].num_token == 8


def test_request_config_is_determined_like_the_reference():
    """GenerationConfig.determined = AsyncEngine._determine_gen_config (async_engine.py:417-437)."""
    tok = _FakeTokenizer()
    g = GenerationConfig(temperature=0.3, top_k=40, repetition_penalty=1.2, n=3, max_new_tokens=None, stop_words=["ld"],
                         stop_token_ids=[9])
    d = g.determined(input_len=10, session_len=64, tokenizer=tok, hf_gen_cfg={"eos_token_id": [1, 14]}, eos_token_id=1)
    assert (d.top_k, d.temperature, d.repetition_penalty, d.n, d.max_new_tokens) == (1, 1.0, 1.0, 1, 54)
    assert d.stop_token_ids == [1, 7, 9, 14]
    assert (g.top_k, g.n, g.max_new_tokens, g.stop_token_ids) == (40, 3, None, [9])          # the user's object is untouched
    d = GenerationConfig(do_sample=True, top_k=1, temperature=0.5).determined(70, 64)
    assert d.temperature == 0.5 and d.max_new_tokens == 512 and d.stop_token_ids == []
    assert GenerationConfig(max_new_tokens=None).determined(70, 64).max_new_tokens == 0
    with pytest.raises(ValueError):
        GenerationConfig(stop_words=["x"]).determined(1, 64)                                  # no tokenizer


def test_pipeline_merges_generation_config_eos_ids(tmp_path):
    """generation_config.json's eos ids stop generation (Llama-3 chat models end turns with an id that is only there)."""
    p = _fake_pipeline()
    base = p.infer([[2, 7]], GenerationConfig(max_new_tokens=9))[0]
    assert base.finish_reason == "length"
    stop = base.token_ids[3]
    p = _fake_pipeline()
    p.hf_gen_cfg = {"eos_token_id": [stop, 99]}
    r = p.infer([[2, 7]], GenerationConfig(max_new_tokens=9))[0]
    assert r.finish_reason == "stop" and r.token_ids == base.token_ids[:base.token_ids.index(stop)]
    r = p.infer([[2, 7]], GenerationConfig(max_new_tokens=9, ignore_eos=True))[0]
    assert r.finish_reason == "length" and r.token_ids == base.token_ids


def test_chat_sessions_carry_the_history():
    """Pipeline.chat (lmdeploy/pipeline.py:206-262): every turn renders history + prompt through the chat template."""
    from lmdeploy_b200.messages import Response, Session
    seen = []

    class Tok(_FakeTokenizer):
        def apply_chat_template(self, messages, add_generation_prompt=True, tokenize=True):
            seen.append([(m["role"], m["content"]) for m in messages])
            out = [0]
            for m in messages:
                out += self.encode(m["content"])
            return out + [1]

    p = _fake_pipeline()
    p.tokenizer = Tok()
    cfg = GenerationConfig(max_new_tokens=3, ignore_eos=True, skip_special_tokens=False)
    s = p.chat("!", gen_config=cfg)
    assert isinstance(s, Session) and len(s.history) == 1 and s.response.finish_reason == "length"
    assert s.history[0] == ("!", s.response.text) and len(s.response.token_ids) == 3
    chunks = list(p.chat("!!", session=s, gen_config=cfg, stream_response=True))
    assert chunks[-1].finish_reason == "length" and len(s.history) == 2
    assert seen[-1] == [("user", "!"), ("assistant", s.history[0][1]), ("user", "!!")]
    assert s.history[1][1] == "".join(c.text for c in chunks)
    # Response.extend merges chunks the way the reference does
    a = Response(text="ab", generate_token_len=2, input_token_len=5, token_ids=[1, 2])
    b = Response(text="c", generate_token_len=3, input_token_len=5, finish_reason="stop", token_ids=[3], index=4)
    m = a.extend(b)
    assert m is a and (m.text, m.token_ids, m.generate_token_len, m.finish_reason, m.index) == ("abc", [1, 2, 3], 3, "stop", 4)
    assert p.session().session_id != p.session().session_id


@pytest.mark.parametrize("prefix_caching", [False, True])
def test_kv_exhaustion_preempts_the_youngest_and_nothing_is_lost(prefix_caching):
    """3 sequences x 20 tokens = 15 blocks of 4, the pool has 9: the youngest running sequence gives its blocks back,
    waits, refills its cache (prompt + what it had generated) and continues.  The KV-content-hashing fake model makes a
    wrong refill visible: every sequence must produce exactly the tokens it produces with an unlimited pool."""
    prompts = [[11, 12, 13, 14, 15, 16], [21, 22, 23, 24, 25, 26], [31, 32, 33, 34, 35, 36]]

    def go(num_blocks):
        sched = Scheduler(max_batch=4, block_len=4, num_blocks=num_blocks, session_len=4096, prefix_caching=prefix_caching)
        model = FakeKvModel(num_blocks, 4)
        seqs = [Sequence(session_id=i, prompt=list(p), max_new_tokens=14, stop_token_ids=[], ignore_eos=True, eos_id=None)
                for i, p in enumerate(prompts)]
        for q in seqs:
            sched.add(q)
        ticks = 0
        while sched.has_work:
            ticks += 1
            assert ticks < 2000
            batch = sched.schedule()
            if not batch:
                continue
            for q in batch:
                assert len(q.blocks) * 4 >= q.cached + 1
            owned = [b for q in sched.active for b in set(q.blocks)]
            if not prefix_caching:
                assert len(owned) == len(set(owned))
            sched.update(batch, model.step(batch))
        return seqs, sched

    free, _ = go(64)
    tight, sched = go(9)
    assert sched.preemptions > 0 and sum(q.preemptions for q in tight) == sched.preemptions
    assert tight[0].preemptions == 0                            # the oldest sequence is never the victim
    for a, b in zip(free, tight):
        assert b.finish_reason == "length" and len(b.generated) == 14 and a.generated == b.generated
    m = sched.metrics()
    assert m["active_seqs"] == 0 and m["waiting_seqs"] == 0 and m["free_blocks"] + m["cached_blocks"] == 9


def test_a_sequence_that_outgrows_the_whole_pool_ends_with_length():
    sched = Scheduler(max_batch=2, block_len=4, num_blocks=3, session_len=4096)
    a = Sequence(session_id=1, prompt=[1, 2, 3], max_new_tokens=50, stop_token_ids=[], ignore_eos=True, eos_id=None)
    sched.add(a)
    run(sched, lambda q: 7)
    # 12 slots: 3 prompt tokens + 9 generated ones are cached, the 10th sampled token finds no block for its K/V
    assert a.finished and a.finish_reason == "length" and len(a.generated) == 10 and a.preemptions == 0
    assert sched.alloc.num_free == 3 and not sched.has_work


def test_dynamic_ntk_prompts_beyond_the_trained_context_are_refused_not_mis_rotated():
    from lmdeploy_b200.turbomind.engine import TurboMindInstance
    from lmdeploy_b200.turbomind.specs import RopeSpec
    import types
    p = _fake_pipeline()
    p.engine.spec = types.SimpleNamespace(rope=RopeSpec(type="dynamic", dim=128, base=1e6, factor=2.0, max_position_embeddings=8))
    p.engine.scheduler.session_len = 64
    inst = TurboMindInstance(p.engine)
    inst._submit(1, list(range(8)), GenerationConfig(max_new_tokens=2))                # fits the trained context
    with pytest.raises(NotImplementedError):
        inst._submit(2, list(range(9)), GenerationConfig(max_new_tokens=2))
    p.engine.spec.rope = RopeSpec(type="dynamic", dim=128, base=1e6, factor=1.0, max_position_embeddings=8)
    inst._submit(3, list(range(9)), GenerationConfig(max_new_tokens=2))                # factor <= 1: the base never changes


def test_engine_config_validation_and_refusals():
    import lmdeploy_b200.messages as M
    for kw in (dict(rope_scaling_factor=-1.0), dict(max_prefill_token_num=-1), dict(num_tokens_per_iter=-2), dict(cache_prompt="x"),
               dict(cache_generation="some"), dict(cache_checkpoint_interval=0), dict(cache_prompt_boundary_skip=0),
               dict(async_=2), dict(dtype="float32"), dict(tp=0)):
        with pytest.raises(AssertionError):
            M.TurbomindEngineConfig(**kw)
    for kw in (dict(empty_init=True), dict(devices=[3]), dict(tp=2, devices=[0, 2])):
        with pytest.raises(NotImplementedError):
            M.TurbomindEngineConfig(**kw)
    M.TurbomindEngineConfig(tp=2, devices=[0, 1], async_=0, use_logn_attn=True, rope_scaling_factor=2.0)     # inert in the reference too
