"""Tensor-parallel host logic on CPU with a real 2-process world (gloo): every rank computes ITS shard of the oracle
decode step from loader-sharded weights, the two per-layer exchanges are real all-reduces, vocab-sharded logits are
all-gathered.  The result must equal the single-process oracle with TP semantics (rank-ordered sum)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import attention as OA
from oracle import linear as OL
from oracle.decode_step import TinyWeights, decode_step
from oracle.numerics import F16, rn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = TinyWeights(hidden=512, layers=2, heads=4, kv_heads=2, inter=512, vocab=256, seed=5)
    sw = w.shard(rank, world)
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e4)
    cache = OA.PagedKvCache(2, w.layers, OA.KvLayout(sw.kv_heads, 64, 128, 8), F16)
    tokens, hist = np.array([3, 9]), np.array([0, 0], np.int32)
    dtype = F16

    def allreduce_T(x):
        """rank-ordered sum in T (what the fused NVLink kernel does): gather all partials, add in rank order"""
        parts = [torch.zeros_like(torch.from_numpy(x)) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(x)))
        acc = parts[0].numpy()
        for p in parts[1:]:
            acc = rn(acc + p.numpy(), dtype)
        return acc

    residual = w.embed[tokens]
    x = OL.rmsnorm(residual, w.L[0]["attn_norm"], 1e-5, dtype)
    for l in range(w.layers):
        lw = sw.L[l]
        qkv = OL.linear(x, lw["w_qkv"]["deq"], dtype)
        attn = OA.decode_attention(qkv, cache, [[0], [1]], hist, l, sw.heads, sw.kv_heads, rope)
        o = allreduce_T(OL.linear(attn, lw["wo"]["deq"], dtype))
        x, residual = OL.residual_rmsnorm(o, residual, w.L[l]["ffn_norm"], 1e-5, dtype)
        a = OL.linear_gated_silu(x, lw["w1w3"]["deq"], dtype)
        f = allreduce_T(OL.linear(a, lw["w2"]["deq"], dtype))
        nxt = w.L[l + 1]["attn_norm"] if l + 1 < w.layers else w.final_norm
        x, residual = OL.residual_rmsnorm(f, residual, nxt, 1e-5, dtype)
    local = torch.from_numpy(OL.linear(x, sw.output["deq"], dtype))
    parts = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(parts, local)
    logits = torch.cat(parts, 1).numpy()
    if rank == 0:
        q.put(logits)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_tp2_gloo_matches_single_process_oracle():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    logits = q.get(timeout=150)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    w = TinyWeights(hidden=512, layers=2, heads=4, kv_heads=2, inter=512, vocab=256, seed=5)
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=128, base=1e4)
    caches = [OA.PagedKvCache(2, w.layers, OA.KvLayout(1, 64, 128, 8), F16) for _ in range(world)]
    ref, _ = decode_step(w, np.array([3, 9]), None, [[0], [1]], np.array([0, 0], np.int32), rope,
                         tp_shards=[w.shard(r, world) for r in range(world)], caches=caches)
    assert (logits == ref).all()          # identical arithmetic, identical order -> bit-identical
    # and TP changes the single-GPU result only by fp16 summation order
    c1 = OA.PagedKvCache(2, w.layers, OA.KvLayout(2, 64, 128, 8), F16)
    ref1, _ = decode_step(w, np.array([3, 9]), c1, [[0], [1]], np.array([0, 0], np.int32), rope)
    assert np.abs(ref1 - ref).max() < 2e-3


# ---------------------------------------------------------------------------------------------------------------
# engine control plane: rank 0 schedules and publishes every step, the other rank follows
# ---------------------------------------------------------------------------------------------------------------
def _engine_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lmdeploy_b200.turbomind.engine import marshal_step
    from lmdeploy_b200.turbomind.scheduler import Scheduler, Sequence
    from lmdeploy_b200.turbomind.tp_engine import StepChannel
    block_len, num_blocks = 4, 32
    store = [[None] * block_len for _ in range(num_blocks)]        # this rank's "KV pool": which token sits where

    def execute(st):
        """What TurboMind._execute does with the arrays, on a toy pool: write the fed tokens at their positions through
        the block table, then 'sample' a hash of everything each sequence can see."""
        out = []
        for i in range(st["batch_size"]):
            rows = st["token_ids"][st["cu_q_len"][i]:st["cu_q_len"][i + 1]]
            k_len = st["cu_k_len"][i + 1] - st["cu_k_len"][i]
            blocks = st["block_ids"][st["cu_block_nums"][i]:st["cu_block_nums"][i + 1]]
            hist = k_len - len(rows)
            for j, t in enumerate(rows):
                store[blocks[(hist + j) // block_len]][(hist + j) % block_len] = t
            h = 1469598103934665603
            for p in range(k_len):
                h = ((h ^ (store[blocks[p // block_len]][p % block_len] + 1)) * 1099511628211) % (1 << 64)
            out += [0] * (len(rows) - 1) + [h % 50000]
        return out

    ch = StepChannel(rank, world)
    steps = 0
    if ch.is_leader:
        sched = Scheduler(max_batch=3, block_len=block_len, num_blocks=num_blocks, session_len=256, prefill_chunk=5,
                          token_budget=8, prefix_caching=True)
        prompts = [list(range(10, 23)), [7], list(range(10, 20)) + [1, 2]]
        seqs = [Sequence(session_id=i, prompt=p, max_new_tokens=4, stop_token_ids=[], ignore_eos=True, eos_id=None)
                for i, p in enumerate(prompts)]
        for s in seqs:
            sched.add(s)
        while sched.has_work:
            batch = sched.schedule()
            st = marshal_step(batch)
            ch.publish(st)
            out = execute(st)
            sched.update(batch, [out[st["cu_q_len"][i + 1] - 1] for i in range(len(batch))])
            steps += 1
        ch.shutdown()
        q.put((rank, steps, [list(s.generated) for s in seqs], store))
    else:
        while True:
            st = ch.receive()
            if st is None:
                break
            execute(st)
            steps += 1
        assert not ch.open
        q.put((rank, steps, None, store))
    dist.destroy_process_group()


def test_engine_leader_publishes_every_step_and_followers_mirror_the_kv_pool():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_engine_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, steps0, generated, store0), (_, steps1, _, store1) = res
    assert steps0 == steps1 > 0
    assert store0 == store1                       # the follower wrote exactly what the leader wrote, where it wrote it
    assert all(len(g) == 4 for g in generated)


# ---------------------------------------------------------------------------------------------------------------
# tp_engine.launch() itself, in two processes, with the device replaced by the dry-run stand-ins (tests/dryrun.py):
# the torchrun entry point a 2-GPU run will execute -- rendezvous, per-rank model shard, KV pool sizing by the
# all-reduced minimum, leader drives generation, followers serve and return when the leader closes.
# ---------------------------------------------------------------------------------------------------------------
def _launch_worker(rank, world, port, q):
    try:
        os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                          MASTER_PORT=str(port))
        import torch
        import torch.distributed as dist
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from lmdeploy_b200 import _lib as L
        import dryrun
        real = L.load()
        rec = dryrun.install(setattr, real)
        import lmdeploy_b200.turbomind.runtime as RT
        import lmdeploy_b200.turbomind.tp as TP
        from lmdeploy_b200.messages import GenerationConfig, TurbomindEngineConfig
        from lmdeploy_b200.turbomind.engine import batched_generate
        from lmdeploy_b200.turbomind.specs import ModelSpec
        from lmdeploy_b200.turbomind import tp_engine

        class CpuModel(RT.DecodeModel):
            def __init__(self, spec, **kw):
                super().__init__(spec, device="cpu", **kw)

        class NoLinks:                       # the NVLink communicator needs peers' device memory: not part of this dry run
            def __init__(self, model, rank, world):
                self.closed = False

            def close(self):
                self.closed = True
        RT.DecodeModel = CpuModel
        TP.TpGroup = NoLinks
        init = dist.init_process_group
        dist.init_process_group = lambda backend, **kw: init("gloo", rank=rank, world_size=world)
        # ranks see different amounts of free memory: the pool must be sized by the smallest
        torch.cuda.mem_get_info = lambda *a, **k: ((3 + rank) << 24, 1 << 30)
        spec = ModelSpec("tiny", "llama", 512, 2, 4, 2, 1024, 1000, 1e4, weight_format="awq")
        cfg = TurbomindEngineConfig(tp=world, max_batch_size=4, session_len=64, cache_max_entry_count=0.5, quant_policy=8)
        eng = tp_engine.launch(spec, cfg)
        steps = sum(rec.names().count(n) for n in ("b200_model_decode_step", "b200_model_chunk_step", "graph.replay"))
        if rank == 0:
            blocks = eng.pool.num_blocks
            seqs = batched_generate(eng, [[1, 2, 3], [4, 5]], GenerationConfig(max_new_tokens=3, ignore_eos=True))
            steps = sum(rec.names().count(n) for n in ("b200_model_decode_step", "b200_model_chunk_step", "graph.replay"))
            eng.close()
            q.put((rank, "ok", [s.generated for s in seqs], steps, blocks, eng._tp is None))
        else:
            q.put((rank, "ok", eng, steps, None, None))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "error", traceback.format_exc(), 0, None, None))


def test_tp_engine_launch_entry_point_in_two_processes():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_launch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        r = q.get(timeout=240)
        out[r[0]] = r
    for p in procs:
        p.join(timeout=60)
    assert out[0][1] == "ok", out[0][2]
    assert out[1][1] == "ok", out[1][2]
    assert out[0][2] == [[0, 0, 0], [0, 0, 0]]                   # the stand-in step leaves id 0 in the output buffer
    assert out[1][2] is None                                     # followers get no engine back
    # every step the leader ran, the follower ran too (chunked prefill: the prompts take one mixed step, then 2-3 decode steps)
    assert out[0][3] == out[1][3] and out[0][3] >= 3
    layer_bytes = 1 * 2 * 64 * (128 + 4) * 2                     # 1 kv head per rank, K|V, 64 tokens, int8 + params; 2 layers
    assert out[0][4] == int((3 << 24) * 0.5) // layer_bytes      # sized by the rank with the least free memory
    assert out[0][5] is True                                     # close() released the communicator


# ---------------------------------------------------------------------------------------------------------------
# bench.py under torchrun (N > 1): rank 0 prints the one JSON line, timing is the maximum over ranks
# ---------------------------------------------------------------------------------------------------------------
def _bench_worker(rank, world, port, q):
    try:
        os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                          MASTER_PORT=str(port))
        import contextlib
        import io
        import types
        import torch
        import torch.distributed as dist
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, os.path.join(root, "tests"))
        sys.path.insert(0, root)
        from lmdeploy_b200 import _lib as L
        import dryrun
        dryrun.install(setattr, L.load())
        import bench
        import lmdeploy_b200.turbomind as TM
        import lmdeploy_b200.turbomind.runtime as RT
        import lmdeploy_b200.turbomind.tp as TP
        from lmdeploy_b200.turbomind.specs import ModelSpec

        class CpuModel(RT.DecodeModel):
            def __init__(self, spec, **kw):
                super().__init__(spec, device="cpu", **kw)

        class NoLinks:
            def __init__(self, *a, **k):
                pass

            def close(self):
                pass

        class Event:
            def __init__(self, *a, **k):
                pass

            def record(self, *a):
                pass

            def elapsed_time(self, other):
                return 2.0 + rank                      # rank 1 is slower: the line must report ITS time
        TM.DecodeModel, TP.TpGroup, torch.cuda.Event = CpuModel, NoLinks, Event
        for name in ("tensor", "zeros", "arange", "randn", "randint", "empty"):
            orig = getattr(torch, name)
            setattr(torch, name, lambda *a, _o=orig, **k: _o(*a, **{**k, "device": "cpu"} if "device" in k else k))
        gen = torch.Generator
        torch.Generator = lambda device=None: gen()
        init = dist.init_process_group
        dist.init_process_group = lambda backend, **kw: init("gloo", rank=rank, world_size=world)
        TM.SPECS["tiny"] = ModelSpec("tiny", "llama", 512, 2, 4, 2, 1024, 1000, 1e4, weight_format="awq")
        bench.CONFIGS["c2"] = ("tiny", 4, 100, 8)
        args = types.SimpleNamespace(gpus=world, steps=4, warmup=3, impl="b200", config="c2", no_graph=False,
                                     no_cpu_baseline=False, attn_iters=2, engine_ingest_limit=150000)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            bench.run_b200(args)
        q.put((rank, "ok", buf.getvalue()))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "error", traceback.format_exc()))


def test_bench_under_torchrun_prints_one_line_with_the_slowest_rank_time():
    import json
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bench_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        r = q.get(timeout=300)
        out[r[0]] = r
    for p in procs:
        p.join(timeout=60)
    assert out[0][1] == "ok", out[0][2]
    assert out[1][1] == "ok", out[1][2]
    lines = [ln for ln in out[0][2].splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and not [ln for ln in out[1][2].splitlines() if ln.startswith("{")]
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["config"]["parallelism"] == "tp2" and d["scaling"] == "strong"
    assert d["ms_per_step"] == pytest.approx(3.0 / 4) and d["value"] == pytest.approx(4 * 4 / 3e-3)      # max over ranks
    assert "cpu_baseline" not in d and d["e2e"]["value"] > 0 and "clocks" in d
