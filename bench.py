#!/usr/bin/env python
"""bench.py -- decode tokens/s of the B200-native TurboMind decode path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2|c3|c1-tiny]

Workload (BASELINE.json configs[1], "C2"): InternLM2-7B geometry, W4A16 AWQ (u4 g128) weights + KV int8, bs = 64,
2K context, greedy decode; synthetic random weights / KV (no checkpoints exist in the image).  A "step" is one
decode forward for the whole batch (64 new tokens).  N > 1 = tensor parallel over N GPUs of one box (heads / FFN /
vocab sharded, fused NVLink all-reduce), launched by torchrun, one rank per GPU: total work is fixed => "strong".

Prints ONE JSON line (rank 0).  Keys follow the driver contract; `roofline` is for the dominant kernel
(decode_attention_kernel), `cpu_baseline` is the oracle port timed on the host cores, `e2e` goes through the
public engine API with host token buffers (H2D / D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu --set full
# capture (profiles/): filled per config once measured
TRAFFIC_NCU = {"c2": 278094848 + 5309696}   # profiles/r02_attention_ncu.md (r02_attn.ncu-rep): dram read + write per launch

CONFIGS = {
    # name: (spec, batch, ctx, kv quant policy)
    "c2": ("internlm2-7b", 64, 2048, 8),
    "c3": ("llama3-8b", 128, 4096, 4),
    "c4": ("internlm2-20b", 64, 2048, 0),
    "c1-tiny": ("internlm2-1.8b", 1, 128, 0),
    # Llama-3-70B FP8 + KV-int8 at tensor-parallel 8 (BASELINE config 5): 70 GB of weights + 44 GB of KV, so it needs
    # --gpus >= 2; launch under torchrun with --gpus 8 for the configuration as named
    "c5": ("llama3-70b", 32, 8192, 8),
}

MIN_GPUS = {"c5": 2}
WEIGHT_LABEL = {"awq": "W4A16", "fp16": "fp16", "fp8": "FP8"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=list(CONFIGS))
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step in a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--attn-iters", type=int, default=20)
    ap.add_argument("--engine-ingest-limit", type=int, default=150000,
                    help="prompt tokens (batch x ctx) up to which the e2e arm ingests the prompts through the engine itself")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _gpu_uuid(torch, local):
    """'GPU-xxxxxxxx-…' of torch device `local` (NVML / nvidia-smi enumerate ALL GPUs; CUDA_VISIBLE_DEVICES may not)."""
    try:
        u = str(torch.cuda.get_device_properties(local).uuid)
        return u if u.startswith("GPU-") else "GPU-" + u
    except Exception:
        return None


class ClockSampler:
    """SM clock + throttle reasons of one GPU, sampled WHILE the timed region runs.

    Primary backend: NVML in this process (pynvml), one sample every 2 ms from a thread, so even a 50 ms timed region
    (TP8) holds tens of samples.  Fallback: `nvidia-smi -lms`, started by start() BEFORE the warm-up because the tool
    takes about a second to come up on an 8-GPU box (an N=8 run of round 1 ended with zero samples that way).  Either
    way only samples stamped inside [begin, end] of the `with` block count."""

    # NVML clocks-event-reason bits (nvml.h; stable ABI)
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
    SMI_REASONS = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0, uuid=None, nvml=None, period_s=0.002):
        self.index, self.uuid, self.period_s = index, uuid, period_s
        self.samples = []             # (t, sm_mhz, reasons set)
        self.sm_max = None
        self.t_begin = self.t_end = None
        self.backend = None
        self._stop = threading.Event()
        self._thread = self._proc = self._nvml = self._handle = None
        self._nvml_module = nvml

    # -- backends ------------------------------------------------------------------------------------------------
    def _open_nvml(self):
        nv = self._nvml_module
        if nv is None:
            import pynvml as nv
        nv.nvmlInit()
        handle = None
        if self.uuid:
            for u in (self.uuid, self.uuid.encode() if isinstance(self.uuid, str) else self.uuid):
                try:
                    handle = nv.nvmlDeviceGetHandleByUUID(u)
                    break
                except Exception:
                    handle = None
        if handle is None:
            handle = nv.nvmlDeviceGetHandleByIndex(self.index)
        self._reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        self.sm_max = int(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
        self._nvml, self._handle = nv, handle
        self._sample_nvml()           # one call up front: a broken query fails here, not in the thread
        self.samples.clear()

    def _sample_nvml(self):
        nv = self._nvml
        mhz = int(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM))
        mask = int(self._reasons_fn(self._handle))
        self.samples.append((time.perf_counter(), mhz, {n for n, b in self.REASON_BITS.items() if mask & b}))

    def _loop_nvml(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml()
            except Exception:
                return
            self._stop.wait(self.period_s)

    def _open_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        sel = self.uuid if self.uuid else str(self.index)
        self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", sel,
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

    def _loop_smi(self):
        for line in self._proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if not r or not r[0].isdigit():
                continue
            if len(r) > 1 and r[1].isdigit():
                self.sm_max = max(self.sm_max or 0, int(r[1]))
            reasons = {n for i, n in enumerate(self.SMI_REASONS) if len(r) > 2 + i and r[2 + i].lower().startswith("active")}
            self.samples.append((time.perf_counter(), int(r[0]), reasons))

    # -- lifecycle -----------------------------------------------------------------------------------------------
    def start(self):
        """Bring the backend up (call before the warm-up).  Idempotent; never raises."""
        if self.backend is not None:
            return self
        try:
            self._open_nvml()
            self.backend, target = "nvml", self._loop_nvml
        except Exception:
            try:
                self._open_smi()
                self.backend, target = "nvidia-smi", self._loop_smi
            except Exception:
                self.backend = "none"
                return self
        self._thread = threading.Thread(target=target, daemon=True)
        self._thread.start()
        return self

    def wait_ready(self, limit_s=5.0):
        """Block (bounded) until the fallback tool printed its first row.  Call BEFORE the barrier that precedes the
        timed region, never inside it: ranks must enter the region together."""
        self.start()
        if self.backend == "nvidia-smi":
            t0 = time.perf_counter()
            while not self.samples and time.perf_counter() - t0 < limit_s and self._proc.poll() is None:
                time.sleep(0.01)
        return self

    def __enter__(self):
        self.start()
        self.t_begin = time.perf_counter()
        return self

    def __exit__(self, *a):
        self.t_end = time.perf_counter()
        if self.backend == "nvidia-smi":
            time.sleep(0.03)                      # one more period: a row that describes the tail of the region
        self._stop.set()
        if self._proc is not None:
            self._proc.terminate()
        if self._thread is not None:
            self._thread.join(timeout=1.0)

    def summary(self):
        slack = 0.03 if self.backend == "nvidia-smi" else 0.0
        inside = [s for s in self.samples
                  if self.t_begin is not None and self.t_begin <= s[0] <= (self.t_end or float("inf")) + slack]
        sm = sorted(s[1] for s in inside)
        reasons = set().union(*[s[2] for s in inside]) if inside else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(reasons),
                "samples": len(sm), "source": self.backend}


# ----------------------------------------------------------------------------------------------------
# reference arm: the CPU port of the path, all host threads (the reference has no CPU engine; see oracle/)
# ----------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lmdeploy_b200.turbomind.specs import SPECS
    from oracle import cpu_baseline
    name, batch, ctx, qp = CONFIGS[args.config]
    spec = SPECS[name]
    cores = cpu_baseline.usable_cores()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_baseline.time_decode(spec, batch, ctx, layers_sample=1, reps=1, threads=cores)
    reps = max(1, min(args.steps, 3))
    r = cpu_baseline.time_decode(spec, batch, ctx, layers_sample=1, reps=reps, threads=cores)
    line = {
        "impl": "reference", "metric": "decode tokens/sec", "value": r["tokens_per_s"], "unit": "tokens/s",
        # `steps` / `warmup` / `ms_per_step` = what actually ran: each "step" of this arm is the bounded sample named in
        # cpu_baseline.sample (one decoder layer + lm_head); `value` extrapolates it to the whole model
        "n_gpus": args.gpus, "steps": reps, "warmup": max(0, min(args.warmup, 1)), "steps_requested": args.steps,
        "warmup_requested": args.warmup, "ms_per_step": (r["t_layer"] + r["t_head"]) * 1e3,
        "ms_per_full_step_extrapolated": r["t_step"] * 1e3,   # what `value` is computed from: layers x t_layer + t_head
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name} {WEIGHT_LABEL.get(spec.weight_format, spec.weight_format)}+"
                               f"{'KV-int%d' % qp if qp else 'KV-fp16'} decode bs={batch} ctx={ctx} (CPU port of the path)"},
        "cpu_baseline": {"value": r["tokens_per_s"], "unit": "tokens/s", "cores": cores, "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["tokens_per_s"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_reference_attention(runner, ref_lib, iters=20, layer=0):
    """Benchmark aid (measurement only): CUDA-event time of the REFERENCE's decode attention -- its main kernel with the
    in-kernel prologue + its reduce, through oracle/_ref -- on the very buffers StaticBatchRunner.time_attention_kernel
    used."""
    import ctypes as C
    import torch
    p = runner._attn_params[0]
    m = runner.model
    ref = ref_lib.load()
    p.debug_flags = 0
    a, keep = ref_lib.args_from_b200(p, torch)
    layers = m.spec.layers
    for i in range(3):
        a.layer_offset = ((layer + i) % layers) * m.layer_bytes
        assert ref.ref_decode_attention(C.byref(a)) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        a.layer_offset = ((layer + i) % layers) * m.layer_bytes
        ref.ref_decode_attention(C.byref(a))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def run_engine_e2e(args, torch, dist, model, spec, batch, ctx, qp, world, barrier):
    """`e2e`: decode tokens/s through the public engine API (TurboMind / TurboMindInstance, the objects pipeline() drives)
    with host token buffers; every step's H2D (control arrays, block table when it changed) and D2H (sampled ids) are
    inside the timed region."""
    import numpy as np
    from lmdeploy_b200.messages import GenerationConfig, TurbomindEngineConfig
    from lmdeploy_b200.turbomind.engine import TurboMind, TurboMindInstance
    steps, warm = args.steps, min(args.warmup, 3)
    session = ctx + warm + steps + 8
    blocks = batch * ((session + model.block_len - 1) // model.block_len) + 4
    cfg = TurbomindEngineConfig(max_batch_size=batch, session_len=session, cache_max_entry_count=blocks, quant_policy=qp,
                                cache_block_seq_len=model.block_len)
    eng = TurboMind(spec, cfg, model=model)
    inst = TurboMindInstance(eng)
    rng = np.random.default_rng(0)
    gen = GenerationConfig(max_new_tokens=warm + steps + 4, ignore_eos=True)
    seqs = [inst._submit(eng.new_session_id(), rng.integers(0, spec.vocab, ctx).tolist(), gen) for _ in range(batch)]
    t_in = time.perf_counter()
    ingest = "engine"
    if batch * ctx > args.engine_ingest_limit:
        # Large contexts (C3: 524k prompt tokens, C5: 262k on 8 GPUs): feeding them through the <= max_batch-row step would
        # take thousands of untimed steps.  The sequences are admitted by the scheduler as usual (blocks, batch slots), then
        # their first ctx - 1 positions are marked cached and filled with synthetic K/V -- as phase A does; the last prompt
        # token and every decode step still go through TurboMind.step().
        ingest = "synthetic KV for the first ctx-1 positions (scheduler-admitted sequences)"
        sched = eng.scheduler
        for s in seqs:
            need = (ctx + model.block_len - 1) // model.block_len
            got = sched.alloc.allocate(need)
            assert got is not None, "KV pool too small for the bench batch"
            s.blocks, s.cached, s.fed = got, ctx - 1, ctx - 1
        sched.waiting.clear()
        sched.active = list(seqs)
        ptrs = torch.tensor([eng.pool.ptr(b) for s in seqs for b in s.blocks], dtype=torch.int64, device=model.device)
        cu_b = torch.arange(batch + 1, dtype=torch.int32, device=model.device) * len(seqs[0].blocks)
        model.prefill_kv_random(eng.pool, ptrs, cu_b, batch, ctx - 1)
    while not all(s.generated for s in seqs):            # prompt ingestion (chunked prefill when the row budget allows)
        eng.step()
    torch.cuda.synchronize()
    ingest_s = time.perf_counter() - t_in
    for _ in range(warm):
        eng.step()
    barrier()
    h2d0, d2h0, tokens = eng.h2d_bytes, eng.d2h_bytes, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        tokens += sum(1 for _, tok in eng.step() if tok is not None)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert tokens == batch * steps, f"engine produced {tokens} tokens in {steps} steps of {batch} sequences"
    out = {"value": tokens / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": (eng.h2d_bytes - h2d0) // steps,
           "d2h_bytes_per_step": (eng.d2h_bytes - d2h0) // steps, "ms_per_step": e2e_s / steps * 1e3,
           "path": "pipeline engine: TurboMindInstance -> TurboMind.step (scheduler, block allocation, pinned staging, "
                   "CUDA graph per batch size)",
           "prompt_ingest": {"how": ingest, "tokens": batch * ctx, "seconds": ingest_s,
                             "tokens_per_s": batch * ctx / ingest_s if ingest == "engine" else None,
                             "steps": eng.scheduler.tick - warm - steps}}
    eng.close()
    return out


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from lmdeploy_b200.turbomind import SPECS, DecodeModel
    from lmdeploy_b200.turbomind.tp import TpGroup
    from lmdeploy_b200.engine import StaticBatchRunner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name, batch, ctx, qp = CONFIGS[args.config]
    spec = SPECS[name]
    if world < MIN_GPUS.get(args.config, 1):
        raise SystemExit(f"config {args.config} does not fit {world} GPU(s): needs --gpus >= {MIN_GPUS[args.config]}")
    total_steps = args.warmup + args.steps
    # two phases run back to back on the same cache: device-resident (value) and host-fed (e2e)
    max_len = ctx + 2 * total_steps + 8

    model = DecodeModel(spec, quant_policy=qp, max_batch=batch, tp_size=world, tp_rank=rank)
    model.load_synthetic(seed=0)
    tp = None
    if world > 1:
        tp = TpGroup(model, rank, world)
    runner = StaticBatchRunner(model, batch=batch, max_len=max_len, seed=0)
    runner.prefill_random(ctx)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- phase A: device-resident steps (value) ----------------
    use_graph = not args.no_graph
    runner.reset_lengths(ctx)
    step_fn = runner.make_device_step(graph=use_graph)
    clocks = ClockSampler(local, uuid=_gpu_uuid(torch, local)).start()
    for _ in range(args.warmup):
        step_fn()
    clocks.wait_ready()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clocks:
        ev0.record()
        for _ in range(args.steps):
            step_fn()
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    barrier()
    ms_per_step = ms / args.steps
    value = batch * args.steps / (ms * 1e-3)

    # ---------------- roofline of the dominant kernel (decode attention), timed alone ----------------
    peak, peak_src = measured_peaks()
    k_attn = ctx + args.warmup + args.steps
    attn = runner.time_attention_kernel(k_attn, iters=args.attn_iters)
    attn_full = runner.time_attention_kernel(k_attn, iters=args.attn_iters, full=True)
    sb = model.step_bytes(batch, ctx + args.warmup + args.steps // 2)
    step_gbs = sb["total"] / (ms_per_step * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "decode_attention_kernel", "achieved": attn["gbs"], "peak": peak, "unit": "GB/s",
        "frac": attn["gbs"] / peak, "traffic": TRAFFIC_NCU.get(args.config), "traffic_source": "ncu --set full capture of this kernel, profiles/r02_attention_ncu.md (not re-measured in this run)", "peak_source": peak_src,
        "bytes_per_launch": attn["bytes"], "us_per_launch": attn["us"],
        # the whole operator (prologue + main kernel + fix-up), what a caller of b200_decode_attention pays
        "full_op": {"us_per_call": attn_full["us"], "achieved": attn_full["gbs"], "frac": attn_full["gbs"] / peak},
        "step": {"algorithmic_bytes": sb["total"], "achieved_gbs_per_gpu": step_gbs, "frac": step_gbs / peak},
    }
    # the kernel to beat: the reference's own decode attention (sm_100a rebuild of its sm80 mma.sync kernels,
    # oracle/_ref) on the same pages, same device, same run -- measurement only, never on the product path.
    # Like for like: the reference call = in-kernel prologue + main kernel + reduce, so it is compared with OUR full op.
    try:
        from oracle import ref_lib
        if ref_lib.available() and world == 1:
            ref_us = time_reference_attention(runner, ref_lib, iters=args.attn_iters)
            roofline["reference_kernel"] = {"what": "TurboMind dispatchDecoding (decoding_sm80_128 + reduce) rebuilt for "
                                                    "sm_100a, same KV pages", "us_per_launch": ref_us,
                                            "achieved": attn["bytes"] / (ref_us * 1e-6) / 1e9,
                                            "speedup_ours": ref_us / attn_full["us"],
                                            "speedup_ours_main_kernel_only": ref_us / attn["us"]}
    except Exception as e:  # the checker library is optional on the bench path
        roofline["reference_kernel"] = {"unavailable": str(e)[:200]}

    # the reference's own tuned W4A16 GEMM (oracle/_ref/libtm_ref_gemm*.so) next to ours on this model's four linears --
    # measurement only (B200_BENCH_REF_GEMM=0 switches it off)
    if os.environ.get("B200_BENCH_REF_GEMM", "0") != "0" and world == 1 and spec.weight_format == "awq":
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import ref_gemm_bench
            D = spec.head_dim
            shapes = [("qkv", spec.hidden, (spec.heads + 2 * spec.kv_heads) * D, 0), ("wo", spec.heads * D, spec.hidden, 0),
                      ("w1w3", spec.hidden, 2 * spec.inter, 1), ("w2", spec.inter, spec.hidden, 0)]
            roofline["reference_gemm"] = ref_gemm_bench.measure(batch, args.attn_iters, shapes, verbose=False)
        except Exception as e:
            roofline["reference_gemm"] = {"unavailable": str(e)[:200]}
    launches = runner.launches_per_step * args.steps

    # ---------------- phase B: end to end through the engine API with host buffers ----------------
    # The continuous-batching engine (TurboMind.step: scheduler, KV block allocation, host marshalling into pinned
    # staging, H2D, CUDA-graph replay per batch size, D2H of the sampled ids, stop checks) drives the same model.  The
    # prompts are ingested through the engine too (outside the timed region: the metric is decode tokens/s at context ctx).
    # Under torchrun every rank runs the same deterministic scheduler on the same requests (SPMD: no per-step broadcast;
    # the sampled ids are identical on every rank by the vocab-sharded argmax exchange).
    del runner
    torch.cuda.empty_cache()
    e2e = run_engine_e2e(args, torch, dist, model, spec, batch, ctx, qp, world, barrier)

    kv_label = f"KV-int{qp}" if qp else "KV-fp16"
    line = {
        "metric": "decode tokens/sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": spec.dtype, "data": "synthetic",
        "config": {"workload": {"awq": f"{name} W4A16(AWQ u4 g128)+{kv_label} greedy decode, bs={batch}, ctx={ctx}",
                                "fp8": f"{name} FP8(e4m3 weights, bf16 activations)+{kv_label} greedy decode, bs={batch}, "
                                       f"ctx={ctx}"}.get(spec.weight_format,
                                                         f"{name} fp16 + KV quant {qp} decode bs={batch} ctx={ctx}"),
                   "parallelism": f"tp{world}", "cuda_graph": use_graph,
                   "l2": "inputs larger than L2 (>= 6 GB of KV + weights streamed per step per GPU)"},
        "roofline": roofline,
        "e2e": e2e,
        "gpu_launches": launches,
        "clocks": clocks.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline
        r = cpu_baseline.time_decode(spec, batch, ctx, layers_sample=1, reps=1, threads=cpu_baseline.usable_cores())
        line["cpu_baseline"] = {"value": r["tokens_per_s"], "unit": "tokens/s", "cores": r["threads"], "kind": "port",
                                "sample": r["sample"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if tp is not None:
            tp.close()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
