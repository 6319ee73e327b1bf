"""CPU baseline of the decode step: the oracle's algorithm restated on torch-CPU tensor ops so it can use every
host core (test/bench infrastructure only -- never on the product path).

BASELINE.json asks for "the reference's PyTorchEngine CPU path"; the reference has NO CPU engine
(lmdeploy/messages.py:531 admits only cuda/ascend/maca/camb -- SURVEY.md section 0), so this is a port ("kind":
"port") structured like PyTorchEngine's default-backend ops:
  * W4A16 linear = dequantise-then-matmul      lmdeploy/pytorch/backends/default/awq_modules.py:37-46
  * RMSNorm                                    lmdeploy/pytorch/backends/default/norm.py:14-27
  * rotary (adjacent pairs after reorder)      lmdeploy/pytorch/backends/default/apply_rotary_emb.py:8-76
  * paged attention: gather pages through the block table, de-quantise (TurboMind KV format), fp32 softmax
Math is fp32 on the CPU (fp16 matmul is not a fast path on x86); results agree with oracle/decode_step.py to
fp32-vs-fp16 rounding, which tests/test_cpu_baseline.py checks on a tiny model.
"""
import math
import time

import torch


def unpack_u4_rows(qweight):
    """int32 [K, N/8] (nibble i = column 8w+i) -> uint8 [K, N]."""
    K, N8 = qweight.shape
    shifts = torch.arange(0, 32, 4, dtype=torch.int32)
    x = (qweight.unsqueeze(-1) >> shifts) & 0xF
    return x.reshape(K, N8 * 8).to(torch.uint8)


def awq_linear(x, qweight, scales, zeros, group=128):
    """x [M,K] fp32; dequant-then-matmul: w = (q - z) * s."""
    q = unpack_u4_rows(qweight).to(torch.float32)
    K, N = q.shape
    w = (q.view(K // group, group, N) - zeros.view(K // group, 1, N)) * scales.view(K // group, 1, N)
    return x @ w.view(K, N)


def rmsnorm(x, w, eps):
    var = x.pow(2).mean(-1, keepdim=True)
    return x * torch.rsqrt(var + eps) * w


def rope_adjacent(x, pos, theta, dim):
    """x [B, H, D]; rotate adjacent pairs."""
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    ang = pos.to(torch.float32)[:, None, None] * inv[None, None, :]
    c, s = torch.cos(ang), torch.sin(ang)
    x0, x1 = x[..., 0:dim:2], x[..., 1:dim:2]
    out = x.clone()
    out[..., 0:dim:2] = c * x0 - s * x1
    out[..., 1:dim:2] = c * x1 + s * x0
    return out


class CpuLayer:
    """One decoder layer with random weights of the real shapes (u4 g128 or dense)."""

    def __init__(self, hidden, heads, kv_heads, inter, fmt, gen, D=128):
        self.hidden, self.heads, self.kv_heads, self.inter, self.fmt, self.D = hidden, heads, kv_heads, inter, fmt, D

        def lin(K, N):
            if fmt == "awq":
                return (torch.randint(-2 ** 31, 2 ** 31 - 1, (K, N // 8), generator=gen, dtype=torch.int32),
                        (torch.rand(K // 128, N, generator=gen) + 0.5) * (0.02 / math.sqrt(K)),
                        torch.randint(6, 11, (K // 128, N), generator=gen).to(torch.float32))
            return (torch.randn(K, N, generator=gen) * (0.1 / math.sqrt(K)),)

        self.wqkv = lin(hidden, (heads + 2 * kv_heads) * D)
        self.wo = lin(heads * D, hidden)
        self.w13 = lin(hidden, 2 * inter)
        self.w2 = lin(inter, hidden)
        self.n1 = torch.ones(hidden)
        self.n2 = torch.ones(hidden)

    def _lin(self, x, w):
        return awq_linear(x, *w) if self.fmt == "awq" else x @ w[0]

    def forward(self, residual, kv, hist, theta, eps=1e-5):
        """kv: dict(kq,vq uint8 [B, Hkv, ctx, D], ks,kz,vs,vz [B,Hkv,ctx]) -- the dequantised-on-the-fly KV of
        this layer (paged gather is modelled by an index_select over a shuffled token axis)."""
        B = residual.shape[0]
        D, Hq, Hkv = self.D, self.heads, self.kv_heads
        x = rmsnorm(residual, self.n1, eps)
        qkv = self._lin(x, self.wqkv)
        q = qkv[:, :Hq * D].view(B, Hq, D)
        k = qkv[:, Hq * D:(Hq + Hkv) * D].view(B, Hkv, D)
        q, k = rope_adjacent(q, hist, theta, D), rope_adjacent(k, hist, theta, D)
        # paged gather + dequant (TurboMind format x = q*scale + zero)
        def gather(q, s, z):
            p = kv["perm"]
            return q.index_select(2, p).to(torch.float32) * s.index_select(2, p)[..., None] + z.index_select(2, p)[..., None]

        K = gather(kv["kq"], kv["ks"], kv["kz"])
        V = gather(kv["vq"], kv["vs"], kv["vz"])
        G = Hq // Hkv
        qg = q.view(B, Hkv, G, D)
        s = torch.einsum("bhgd,bhtd->bhgt", qg, K) / math.sqrt(D)
        p = torch.softmax(s, -1)
        attn = torch.einsum("bhgt,bhtd->bhgd", p, V).reshape(B, Hq * D)
        residual = residual + self._lin(attn, self.wo)
        x = rmsnorm(residual, self.n2, eps)
        gu = self._lin(x, self.w13)
        act = torch.nn.functional.silu(gu[:, 0::2]) * gu[:, 1::2]
        residual = residual + self._lin(act, self.w2)
        return residual


def usable_cores(cap=32):
    """Host threads that can actually run: min(affinity mask, cgroup CPU quota), capped (torch-CPU ops of this size
    stop scaling -- and collapse under oversubscription -- long before 128 threads)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, min(n, cap))


def time_decode(spec, batch, ctx, kv_quant=True, layers_sample=1, reps=1, threads=None, seed=0):
    """Time `layers_sample` decoder layers + lm_head of `spec` at (batch, ctx) on the host cores.
    Returns dict(tokens_per_s (full-model extrapolation), t_layer, t_head, threads, sample)."""
    threads = threads or usable_cores()
    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(seed)
    D = 128
    layer = CpuLayer(spec.hidden, spec.heads, spec.kv_heads, spec.inter, "awq" if spec.weight_format == "awq" else "fp16",
                     gen, D)
    kv = dict(perm=torch.randperm(ctx, generator=gen))
    for n in ("kq", "vq"):
        kv[n] = torch.randint(0, 256, (batch, spec.kv_heads, ctx, D), generator=gen, dtype=torch.uint8)
    for n in ("ks", "vs"):
        kv[n] = torch.rand(batch, spec.kv_heads, ctx, generator=gen) * 0.03
    for n in ("kz", "vz"):
        kv[n] = -torch.rand(batch, spec.kv_heads, ctx, generator=gen) * 4
    head = torch.randn(spec.hidden, spec.vocab, generator=gen) * 0.01
    residual = torch.randn(batch, spec.hidden, generator=gen)
    hist = torch.full((batch,), ctx, dtype=torch.int64)
    t_layer, t_head = [], []
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            r = residual
            for _ in range(layers_sample):
                r = layer.forward(r, kv, hist, spec.rope_theta)
            t1 = time.perf_counter()
            logits = rmsnorm(r, layer.n1, 1e-5) @ head
            tok = logits.argmax(-1)
            t2 = time.perf_counter()
            t_layer.append((t1 - t0) / layers_sample)
            t_head.append(t2 - t1)
    tl, th = min(t_layer), min(t_head)
    step = tl * spec.layers + th
    return dict(tokens_per_s=batch / step, t_layer=tl, t_head=th, t_step=step, threads=threads,
                sample=f"{layers_sample} of {spec.layers} decoder layers + lm_head at bs={batch}, ctx={ctx} "
                       f"(fp32 torch-CPU, weights dequantised per call), extrapolated x{spec.layers}/{layers_sample}",
                checksum=int(tok.sum()))
