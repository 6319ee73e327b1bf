"""Continuous-batching scheduler + KV block accounting (pure host logic, unit-tested on CPU).

Plays the role of Engine::Impl::InternalThreadEntry / Scheduler::Schedule in the reference
(src/turbomind/engine/engine.cc:770-920, engine/scheduler.cc:1018): every step the active batch is re-formed --
waiting requests are admitted while batch slots and KV blocks are available, a sequence gets a new block when it
crosses a block boundary, finished sequences are retired and their blocks returned.

Prompt tokens are consumed through the same step as decode tokens.  By default one prompt token per step
("prefill-by-decode": the decode kernel appends K/V for the token it is given, so feeding prompt tokens reproduces
causal prefill exactly).  With `prefill_chunk` > 1 a sequence still in its prompt feeds up to that many tokens in one
step, mixed with the decoding sequences of the batch (the reference's chunked prefill / "dynamic split-fuse",
unified_attention_layer.cc:310, 613-651); the step's token count never exceeds `token_budget` -- the row capacity of
the device step (b200_model_decode_step handles <= 128 rows).

Prefix caching (TurbomindEngineConfig.enable_prefix_caching, lmdeploy/messages.py:238-239; the reference indexes
logical blocks in src/turbomind/engine/prefix_trie.h): K/V of position p depends only on tokens [0, p], so a FULL block
is identified by (identity of the preceding block, its block_len tokens).  Full blocks are published to an exact
index when they fill; a new request re-uses the longest chain of published blocks covering its prompt (always leaving
at least the last prompt token to be fed, so the step that produces its first logits still runs) and starts with
`cached` already advanced -- those tokens are never fed.  Shared blocks are read-only by construction: a decode step
writes only position `cached`, which lies in the first private block.  Unreferenced published blocks stay resident
and are evicted oldest-first only when the free list runs dry.

Out of KV memory: the youngest running sequence is preempted -- its blocks go back to the pool, it returns to the head
of the waiting queue and, once its whole context fits again, refills its cache (prompt and the tokens it had generated)
and continues.  No request ends early for lack of memory unless it alone outgrows the pool."""
from collections import OrderedDict, deque
from dataclasses import dataclass, field
from typing import Deque, Dict, List, Optional


@dataclass(eq=False)       # identity semantics: sequences are compared and looked up as objects, never by value
class Sequence:
    session_id: int
    prompt: List[int]
    max_new_tokens: int
    stop_token_ids: List[int]
    ignore_eos: bool
    eos_id: Optional[int]
    blocks: List[int] = field(default_factory=list)
    cached: int = 0               # tokens whose K/V are in the cache
    fed: int = 0                  # prompt tokens already fed
    generated: List[int] = field(default_factory=list)
    finished: bool = False
    finish_reason: Optional[str] = None
    reused_tokens: int = 0        # prompt tokens served from the prefix cache (never fed)
    step_tokens: int = 1          # tokens this sequence feeds in the step being formed / executed
    error: Optional[str] = None   # why the request was rejected (finish_reason 'error')
    preemptions: int = 0          # times this sequence lost its blocks to an older one and had to refill its cache
    want_logits: bool = False     # GenerationConfig.output_logits == 'generation': keep the sampled row's logits per token
    want_logprobs: int = 0        # GenerationConfig.logprobs: keep that many top log-probabilities per token
    logits: List = field(default_factory=list)
    logprobs: List = field(default_factory=list)
    chain: List[int] = field(default_factory=list)      # prefix-cache identity of each full block so far
    sampling: Optional[dict] = None   # do_sample requests: temperature / top_k / top_p / min_p / repetition_penalty / seed
    rng_slot: int = -1                # ... and the device RNG state slot the engine gave it

    def token_at(self, pos):
        return self.prompt[pos] if pos < len(self.prompt) else self.generated[pos - len(self.prompt)]

    def next_input(self):
        return self.token_at(self.cached)

    @property
    def context_len(self):
        """Tokens known so far: the prompt and what has been generated."""
        return len(self.prompt) + len(self.generated)

    def next_inputs(self):
        """The `step_tokens` tokens of this step: positions [cached, cached + step_tokens) of prompt + generated.  One
        token (the last generated) in steady decode; several only while the cache is being (re)filled -- inside the
        prompt, or after a preemption dropped this sequence's blocks."""
        return [self.token_at(p) for p in range(self.cached, min(self.cached + self.step_tokens, self.context_len))]

    @property
    def in_prefill(self):
        return self.fed < len(self.prompt) - 1


class InputLengthError(ValueError):
    """The prompt alone fills the session (the reference answers ResponseType.INPUT_LENGTH_ERROR)."""


class BlockAllocator:
    def __init__(self, num_blocks):
        self.free = list(range(num_blocks - 1, -1, -1))
        self.total = num_blocks
        self.evict = None             # optional callable(n) -> blocks reclaimed from a cache layered on top

    def allocate(self, n):
        if n > len(self.free) and self.evict is not None:
            self.free.extend(self.evict(n - len(self.free)))
        if n > len(self.free):
            return None
        return [self.free.pop() for _ in range(n)]

    def release(self, blocks):
        self.free.extend(reversed(blocks))

    @property
    def num_free(self):
        return len(self.free)


class PrefixCache:
    """Exact index of full KV blocks.  A published block has a node id; its key is (parent node id, tokens)."""
    ROOT = 0

    def __init__(self, block_len):
        self.block_len = block_len
        self.index: Dict[tuple, int] = {}            # (parent node, tokens) -> node
        self.block_of: Dict[int, int] = {}           # node -> physical block
        self.node_of: Dict[int, int] = {}            # physical block -> node
        self.key_of: Dict[int, tuple] = {}           # node -> key
        self.refs: Dict[int, int] = {}               # physical block -> live sequences holding it
        self.idle: "OrderedDict[int, None]" = OrderedDict()   # published, unreferenced blocks, oldest first
        self._next_node = 1
        self.hits = self.queries = 0                 # in tokens

    def match(self, tokens, limit):
        """Longest chain of published blocks covering tokens[:limit] -> (blocks, nodes); takes a reference on each."""
        blocks, nodes, parent = [], [], self.ROOT
        bl = self.block_len
        self.queries += len(tokens)
        for j in range(limit // bl):
            node = self.index.get((parent, tuple(tokens[j * bl:(j + 1) * bl])))
            if node is None:
                break
            b = self.block_of[node]
            self.refs[b] = self.refs.get(b, 0) + 1
            self.idle.pop(b, None)
            blocks.append(b)
            nodes.append(node)
            parent = node
        self.hits += len(blocks) * bl
        return blocks, nodes

    def publish(self, parent, tokens, block):
        """Index `block` (just filled with `tokens` after chain `parent`).  Returns its node id; if identical content is
        already indexed (two requests raced through the same prefix) the existing node is returned and `block` stays
        private to its owner."""
        key = (parent, tuple(tokens))
        node = self.index.get(key)
        if node is not None:
            return node
        node = self._next_node
        self._next_node += 1
        self.index[key] = node
        self.block_of[node] = block
        self.node_of[block] = node
        self.key_of[node] = key
        self.refs[block] = 1
        return node

    def release(self, blocks):
        """Drop one reference per block; returns the blocks that are NOT cached (to go back to the free list)."""
        private = []
        for b in blocks:
            if b in self.node_of:
                self.refs[b] -= 1
                if self.refs[b] == 0:
                    self.idle[b] = None
            else:
                private.append(b)
        return private

    def evict(self, n):
        """Forget up to n unreferenced blocks, oldest first.  Descendants of a forgotten node become unreachable (their
        keys name a node id that no longer resolves) and age out the same way."""
        out = []
        while self.idle and len(out) < n:
            b, _ = self.idle.popitem(last=False)
            node = self.node_of.pop(b)
            del self.index[self.key_of.pop(node)]
            del self.block_of[node]
            del self.refs[b]
            out.append(b)
        return out

    @property
    def num_idle(self):
        return len(self.idle)


class Scheduler:
    def __init__(self, *, max_batch, block_len, num_blocks, session_len, prefix_caching=False, prefill_chunk=1,
                 token_budget=None):
        self.max_batch, self.block_len, self.session_len = max_batch, block_len, session_len
        self.prefill_chunk = max(1, int(prefill_chunk))
        self.token_budget = max(max_batch, token_budget or max_batch)
        self.alloc = BlockAllocator(num_blocks)
        self.prefix = PrefixCache(block_len) if prefix_caching else None
        if self.prefix is not None:
            self.alloc.evict = self.prefix.evict
        self.waiting: Deque[Sequence] = deque()
        self.active: List[Sequence] = []
        self.sessions: Dict[int, Sequence] = {}
        self.tick = 0                 # schedule() calls so far
        self.preemptions = 0          # sequences sent back to the waiting queue for lack of KV blocks

    def add(self, seq: Sequence):
        if seq.session_id in self.sessions:
            raise KeyError(f"session {seq.session_id} already exists")
        if len(seq.prompt) == 0:
            raise ValueError("empty prompt")
        if len(seq.prompt) >= self.session_len:
            raise InputLengthError(f"prompt of {len(seq.prompt)} tokens does not fit session_len {self.session_len}")
        if len(seq.prompt) + seq.max_new_tokens > self.session_len:
            seq.max_new_tokens = max(0, self.session_len - len(seq.prompt))
        self.sessions[seq.session_id] = seq
        self.waiting.append(seq)

    def cancel(self, session_id):
        seq = self.sessions.get(session_id)
        if seq and not seq.finished:
            seq.finished, seq.finish_reason = True, "cancel"

    def _release(self, seq):
        blocks = seq.blocks if self.prefix is None else self.prefix.release(seq.blocks)
        self.alloc.release(blocks)
        seq.blocks = []

    def _retire(self):
        """Give back the blocks and the session slot of every finished active sequence."""
        if not any(s.finished for s in self.active):       # the common tick: nothing to give back
            return
        still = []
        for s in self.active:
            if s.finished:
                self._release(s)
                self.sessions.pop(s.session_id, None)
            else:
                still.append(s)
        self.active = still

    def _preempt(self, seq):
        self._release(seq)
        seq.chain, seq.cached, seq.fed, seq.step_tokens = [], 0, 0, 1
        seq.preemptions += 1
        self.preemptions += 1
        self.active.remove(seq)
        self.waiting.appendleft(seq)

    def _blocks_needed(self, seq, upto_tokens):
        return (upto_tokens + self.block_len - 1) // self.block_len - len(seq.blocks)

    def schedule(self):
        """Form the batch of this step.  Returns the sequences that run; each feeds `step_tokens` tokens (1 unless it
        is inside its prompt and prefill_chunk > 1)."""
        self.tick += 1
        self._retire()
        # grow active sequences that cross a block boundary (need room for token index `cached`)
        # Out of KV memory: the YOUNGEST running sequence gives its blocks back and returns to the head of the waiting
        # queue; it refills its cache (prompt + what it generated so far) when it is admitted again and continues -- the
        # recompute flavour of the reference's preemption (SequenceManager frees the blocks of lower-priority sequences,
        # src/turbomind/models/llama/SequenceManager.cc; there is no host swap on this path).  Published prefix blocks
        # stay in the cache index, so with prefix caching a resumed sequence usually finds most of itself again.
        preempted = False
        bl = self.block_len
        for s in list(self.active):
            if s.finished:
                continue
            need = (s.cached + bl) // bl - len(s.blocks)          # == _blocks_needed(s, s.cached + 1)
            if need <= 0:
                continue                   # the common case, decided without the O(n) membership test below
            if s not in self.active:       # preempted earlier in this loop (its blocks are gone, so need > 0 brought us here)
                continue
            while need > 0:
                got = self.alloc.allocate(need)
                if got is not None:
                    s.blocks += got
                    break
                victim = next((v for v in reversed(self.active) if not v.finished), None)
                if victim is None or (victim is s and sum(not v.finished for v in self.active) == 1):
                    # alone and still too long for the whole cache: nothing to wait for
                    s.finished, s.finish_reason = True, "length"
                    break
                self._preempt(victim)
                preempted = True
                if victim is s:
                    break
        self._retire()                     # a sequence stopped just above gives its blocks to this step's admissions
        # admit (not in a step that had to preempt: the freed blocks belong to the sequences that stay)
        while self.waiting and len(self.active) < self.max_batch and not preempted:
            s = self.waiting[0]
            if s.finished:                 # cancelled while waiting
                self.waiting.popleft()
                self.sessions.pop(s.session_id, None)
                continue
            if self.prefix is not None and not s.blocks:
                # never the whole context: its last token must be fed to produce the next logits
                ctx = s.prompt + s.generated if s.generated else s.prompt
                s.blocks, s.chain = self.prefix.match(ctx, len(ctx) - 1)
                s.cached = len(s.blocks) * self.block_len
                s.fed = min(len(s.prompt), s.cached)
                s.reused_tokens = max(s.reused_tokens, s.fed)
            # a resumed sequence comes back only when its whole context fits again (else it would be preempted next step)
            upto = s.context_len + 1 if s.preemptions else s.cached + 1
            if (upto + self.block_len - 1) // self.block_len > self.alloc.total:
                # cannot fit even an empty cache: waiting would never end
                self._release(s)
                s.finished, s.finish_reason = True, "length"
                self.waiting.popleft()
                self.sessions.pop(s.session_id, None)
                continue
            got = self.alloc.allocate(self._blocks_needed(s, upto))
            if got is None:
                if self.prefix is not None:            # give the references back; retry from scratch next step
                    self._release(s)
                    s.chain, s.cached, s.fed = [], 0, 0
                break
            s.blocks += got
            self.waiting.popleft()
            self.active.append(s)
        batch = [s for s in self.active if not s.finished]
        budget = self.token_budget - len(batch)
        chunked = self.prefill_chunk > 1
        for s in batch:
            s.step_tokens = 1
            if not chunked:
                continue
            rem = len(s.prompt) + len(s.generated) - s.cached   # tokens still to feed, the one that produces the next logits included
            if rem > 1 and budget > 0:
                extra = min(self.prefill_chunk - 1, rem - 1, budget)
                need = self._blocks_needed(s, s.cached + 1 + extra)
                got = self.alloc.allocate(need) if need > 0 else []
                if got is None:
                    continue               # no room for a longer chunk now: one token this step
                s.blocks += got
                s.step_tokens += extra
                budget -= extra
        return batch

    def update(self, batch: List[Sequence], sampled: List[int]):
        """Account for one executed step: every sequence cached its `step_tokens` tokens; `sampled[i]` is the token
        sampled from the LAST row of sequence i, emitted once the sequence is past its prompt."""
        events = []
        prefix, bl, session_len = self.prefix, self.block_len, self.session_len
        for s, tok in zip(batch, sampled):
            n = s.step_tokens
            before = s.cached
            cached = s.cached = before + n
            plen = len(s.prompt)
            s.fed = min(plen, s.fed + n)
            s.step_tokens = 1
            if prefix is not None:
                for j in range(before // bl, cached // bl):       # blocks that just filled
                    if len(s.chain) == j:
                        toks = [s.token_at(p) for p in range(j * bl, (j + 1) * bl)]
                        s.chain.append(prefix.publish(s.chain[-1] if s.chain else PrefixCache.ROOT, toks, s.blocks[j]))
            gen = s.generated
            if cached < plen + len(gen):
                continue                   # still filling the cache (prompt, or refill after a preemption): sample discarded
            if s.max_new_tokens == 0:
                s.finished, s.finish_reason = True, "length"
                events.append((s, None))
                continue
            tok = int(tok)
            gen.append(tok)
            # ignore_eos switches off EVERY stop id, user stop words included: the reference only hands stop_ids to its
            # engine when ignore_eos is off (lmdeploy/turbomind/turbomind.py `_get_generation_config`); the stop token
            # itself is appended to the sequence, the pipeline layer hides it from the response
            if not s.ignore_eos and ((s.eos_id is not None and tok == s.eos_id) or tok in s.stop_token_ids):
                s.finished, s.finish_reason = True, "stop"
            elif len(gen) >= s.max_new_tokens or cached + 1 > session_len:
                s.finished, s.finish_reason = True, "length"
            events.append((s, tok))
        self._retire()                     # an idle engine holds no blocks of finished requests
        return events

    @property
    def has_work(self):
        return bool(self.waiting) or any(not s.finished for s in self.active)

    def metrics(self):
        """Counterpart of TurboMind.get_schedule_metrics -> ScheduleMetrics (src/turbomind/python/bind.cpp:530-537, 990-994):
        same field names, plus the raw block counts they derive from."""
        idle = self.prefix.num_idle if self.prefix is not None else 0
        used = self.alloc.total - self.alloc.num_free - idle
        hit = self.prefix.hits / self.prefix.queries if self.prefix is not None and self.prefix.queries else 0.0
        return dict(total_seqs=len(self.sessions), active_seqs=sum(not s.finished for s in self.active),
                    waiting_seqs=sum(not s.finished for s in self.waiting),
                    cache_usage=used / self.alloc.total if self.alloc.total else 0.0, prefix_cache_hit_rate=hit,
                    scheduler_tick=self.tick, preemptions=self.preemptions, total_blocks=self.alloc.total,
                    free_blocks=self.alloc.num_free,
                    cached_blocks=idle, active_blocks=used)
