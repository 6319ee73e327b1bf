"""Continuous-batching scheduler + KV block accounting (pure host logic, unit-tested on CPU).

Plays the role of Engine::Impl::InternalThreadEntry / Scheduler::Schedule in the reference
(src/turbomind/engine/engine.cc:770-920, engine/scheduler.cc:1018): every step the active batch is re-formed --
waiting requests are admitted while batch slots and KV blocks are available, a sequence gets a new block when it
crosses a block boundary, finished sequences are retired and their blocks returned.

Prompt tokens are consumed through the same decode step one token per step ("prefill-by-decode"): the decode
kernel appends K/V for the token it is given, so feeding prompt tokens reproduces causal prefill exactly.  A tiled
prefill attention kernel is SURVEY section 8f rank 1 ("next"), not part of this hot path."""
from collections import deque
from dataclasses import dataclass, field
from typing import Deque, Dict, List, Optional


@dataclass
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

    def next_input(self):
        if self.fed < len(self.prompt):
            return self.prompt[self.fed]
        return self.generated[-1]

    @property
    def in_prefill(self):
        return self.fed < len(self.prompt) - 1


class BlockAllocator:
    def __init__(self, num_blocks):
        self.free = list(range(num_blocks - 1, -1, -1))
        self.total = num_blocks

    def allocate(self, n):
        if n > len(self.free):
            return None
        return [self.free.pop() for _ in range(n)]

    def release(self, blocks):
        self.free.extend(reversed(blocks))

    @property
    def num_free(self):
        return len(self.free)


class Scheduler:
    def __init__(self, *, max_batch, block_len, num_blocks, session_len):
        self.max_batch, self.block_len, self.session_len = max_batch, block_len, session_len
        self.alloc = BlockAllocator(num_blocks)
        self.waiting: Deque[Sequence] = deque()
        self.active: List[Sequence] = []
        self.sessions: Dict[int, Sequence] = {}

    def add(self, seq: Sequence):
        if seq.session_id in self.sessions:
            raise KeyError(f"session {seq.session_id} already exists")
        if len(seq.prompt) == 0:
            raise ValueError("empty prompt")
        if len(seq.prompt) + seq.max_new_tokens > self.session_len:
            seq.max_new_tokens = max(0, self.session_len - len(seq.prompt))
        self.sessions[seq.session_id] = seq
        self.waiting.append(seq)

    def cancel(self, session_id):
        seq = self.sessions.get(session_id)
        if seq and not seq.finished:
            seq.finished, seq.finish_reason = True, "cancel"

    def _blocks_needed(self, seq, upto_tokens):
        return (upto_tokens + self.block_len - 1) // self.block_len - len(seq.blocks)

    def schedule(self):
        """Form the batch of this step.  Returns the list of sequences that run (each feeds one token)."""
        # retire
        still = []
        for s in self.active:
            if s.finished:
                self.alloc.release(s.blocks)
                s.blocks = []
                self.sessions.pop(s.session_id, None)
            else:
                still.append(s)
        self.active = still
        # grow active sequences that cross a block boundary (need room for token index `cached`)
        for s in self.active:
            need = self._blocks_needed(s, s.cached + 1)
            if need > 0:
                got = self.alloc.allocate(need)
                if got is None:            # out of KV memory: stop the youngest sequence (no swap on this path)
                    s.finished, s.finish_reason = True, "length"
                    continue
                s.blocks += got
        # admit
        while self.waiting and len(self.active) < self.max_batch:
            s = self.waiting[0]
            if s.finished:                 # cancelled while waiting
                self.waiting.popleft()
                self.sessions.pop(s.session_id, None)
                continue
            got = self.alloc.allocate(self._blocks_needed(s, 1))
            if got is None:
                break
            s.blocks += got
            self.waiting.popleft()
            self.active.append(s)
        return [s for s in self.active if not s.finished]

    def update(self, batch: List[Sequence], sampled: List[int]):
        """Account for one executed step: every sequence cached one more token; sequences past their prompt
        emit the sampled token."""
        events = []
        for s, tok in zip(batch, sampled):
            s.cached += 1
            if s.fed < len(s.prompt):
                s.fed += 1
            if s.fed < len(s.prompt):
                continue                   # still consuming the prompt; sampled token is discarded
            if s.max_new_tokens == 0:
                s.finished, s.finish_reason = True, "length"
                events.append((s, None))
                continue
            s.generated.append(int(tok))
            stop = (not s.ignore_eos and s.eos_id is not None and tok == s.eos_id) or (tok in s.stop_token_ids)
            if stop:
                s.finished, s.finish_reason = True, "stop"
            elif len(s.generated) >= s.max_new_tokens or s.cached + 1 > self.session_len:
                s.finished, s.finish_reason = True, "length"
            events.append((s, int(tok)))
        return events

    @property
    def has_work(self):
        return bool(self.waiting) or any(not s.finished for s in self.active)
