"""TEST INFRASTRUCTURE -- CPU restatement of the reference's sampling stage for do_sample requests.  Only tests/ may import it.

Follows (file:line in /root/reference/src/turbomind):
  * RepetitionPenaltyKernel        kernels/sampling_penalty_kernels.cu:135-170   logit < 0 ? logit * p : logit / p on seen ids
  * batchApplyTemperaturePenalty_v2 kernels/sampling_penalty_kernels.cu:27-80    logit * (1 / (T + 1e-6))        (fp32)
  * topKSortStage1 / Stage2        kernels/sampling_topk_kernels.cu:97-215       k largest, descending; p = exp(l - l0) / sum,
                                                                                the sum taken sequentially in sorted order
  * invokeSoftmax + topPSort       generation/sampling.cc:126-144                top_k == 0: softmax over the vocabulary, sorted
  * topPMinPFilter                 kernels/sampling_topp_kernels.cu:290-383      prefix until cum > top_p; then q < q0 * min_p out;
                                                                                renormalise by the kept mass
  * sampling                       kernels/sampling_kernels.cu:15-57             first inclusive prefix sum > u, else the last
  * curand_init(seed, 0, 0) / curand_uniform (XORWOW)   curand_kernel.h (CUDA toolkit): InitializeRandomStates,
                                                         kernels/sampling_topk_kernels.cu:74-95

All arithmetic in numpy float32 in the reference's order.  Sorted order on EQUAL values is unspecified in the reference (cub
reduction / radix sort); here equal values are ordered by ascending token id, which is what the device kernel does.
Parity: pinned on a B200 against lmdeploy_b200's kernel (tests/test_sampling_gpu.py: same candidate count, same token for the
same uniform number, XORWOW sequence identical); the reference's own sampling kernels are not built into oracle/_ref
(parity vs the reference binary unpinned for this stage -- its test suite holds no vectors for it either)."""
import numpy as np

F = np.float32


class Xorwow:
    """curandStateXORWOW_t after curand_init(seed, 0, 0); uniform() = curand_uniform()."""

    def __init__(self, seed):
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        m = 0xFFFFFFFF
        s0 = (seed & m) ^ 0xaad26b49
        s1 = ((seed >> 32) & m) ^ 0xf7dcefdd
        t0 = (1099087573 * s0) & m
        t1 = (2591861531 * s1) & m
        self.d = (6615241 + t1 + t0) & m
        self.v = [(123456789 + t0) & m, (362436069 ^ t0) & m, (521288629 + t1) & m, (88675123 ^ t1) & m, (5783321 + t0) & m]

    def next_u32(self):
        m = 0xFFFFFFFF
        v = self.v
        t = (v[0] ^ (v[0] >> 2)) & m
        v[0], v[1], v[2], v[3] = v[1], v[2], v[3], v[4]
        v[4] = ((v[4] ^ ((v[4] << 4) & m)) ^ (t ^ ((t << 1) & m))) & m
        self.d = (self.d + 362437) & m
        return (v[4] + self.d) & m

    def uniform(self):
        # _curand_uniform: x * 2^-32 + 2^-33 in float (curand_uniform.h:69-72)
        x = F(self.next_u32())
        return F(x * F(2.3283064e-10) + F(2.3283064e-10 / 2.0))


def process_logits(logits, temperature=1.0, repetition_penalty=1.0, history=None):
    x = np.asarray(logits, dtype=F).copy()
    if repetition_penalty != 1.0 and history is not None and len(history):
        ids = np.unique(np.asarray(history, dtype=np.int64))
        ids = ids[(ids >= 0) & (ids < x.size)]
        pen = F(repetition_penalty)
        x[ids] = np.where(x[ids] < 0, x[ids] * pen, x[ids] / pen).astype(F)
    scale = F(1.0) / F(F(temperature) + F(1e-6))
    return (x * scale).astype(F)


def candidates(x, top_k=0, top_p=1.0, min_p=0.0):
    """-> (token ids of the final candidate list in the reference's sorted order, their renormalised probabilities)."""
    x = np.asarray(x, dtype=F)
    V = x.size
    order = np.lexsort((np.arange(V), -x.astype(np.float64)))      # value descending, id ascending on ties
    k = int(top_k)
    if k <= 0 or k > V:
        k = 0
    k = min(k, 1024)
    if k > 0:
        ids = order[:k]
        e = np.exp((x[ids] - x[ids[0]]).astype(F)).astype(F)
        s = F(0)
        for v in e:                                                 # sequential sum in sorted order (Stage2, one thread)
            s = F(s + v)
        p = (e / s).astype(F)
    else:
        ids = order
        e = np.exp((x - x.max()).astype(F)).astype(F)
        p = (e / F(e.astype(np.float64).sum())).astype(F)[ids]
    n, total = len(ids), F(1.0)
    if top_p != 1.0:
        cum = np.cumsum(p.astype(np.float64))
        hit = np.nonzero(cum > float(F(top_p)))[0]
        n = int(hit[0]) + 1 if hit.size else len(ids)
        total = F(cum[n - 1])
    if min_p != 0.0:
        eps = F(1e-6)
        q = (p[:n] / F(total + eps)).astype(F)
        cut = F(q[0] * F(min_p))
        below = np.nonzero(q < cut)[0]
        m = int(below[0]) if below.size else n
        m = max(m, 1)
        total = F(total * F(q[:m].astype(np.float64).sum()))
        n = m
    return ids[:n], (p[:n] / total).astype(F)


def draw(ids, probs, u, vocabulary_order=False):
    """First inclusive prefix sum > u (else the last candidate).  vocabulary_order: walk the candidates by ascending token id
    (the device kernel's top_k == 0 path) instead of by descending probability."""
    ids, probs = np.asarray(ids), np.asarray(probs, dtype=np.float64)
    if vocabulary_order:
        o = np.argsort(ids, kind="stable")
        ids, probs = ids[o], probs[o]
    cum = np.cumsum(probs)
    hit = np.nonzero(cum > float(u))[0]
    return int(ids[hit[0]] if hit.size else ids[-1])


def sample(logits, u, *, temperature=1.0, top_k=0, top_p=1.0, min_p=0.0, repetition_penalty=1.0, history=None):
    x = process_logits(logits, temperature, repetition_penalty, history)
    ids, probs = candidates(x, top_k, top_p, min_p)
    return draw(ids, probs, u, vocabulary_order=(top_k <= 0 or top_k > x.size)), len(ids)
