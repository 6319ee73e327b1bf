"""Index-level emulation of csrc/prefill_attention.cuh on the CPU, checked against the independent oracle.

The kernel has not run on a device yet; what CAN be checked without one is everything that is not
tensor-core semantics: the CTA -> (sequence, query tile) map, the row -> (token, head) stacking, page addressing (data and
parameter regions, int4 nibble order), the operand-tile byte addresses (Q [d/8][row][8], K [d/8][key][8], V transposed
[key/8][d][8] with its byte_perm selectors), tile count, causal mask, dead-key zeroing and the online-softmax recurrence.
The emulation below follows the CUDA source statement by statement with the same names; the two tcgen05.mma groups are
replaced by a generic reader of the UMMA canonical K-major layout (element (row, k) at (k/8)*LBO + (row/8)*SBO +
(row%8)*16 + (k%8)*2 -- the layout tools/umma_probe.cu verified on hardware) followed by a float64 matmul."""
import numpy as np
import pytest

from oracle import attention as OA
from oracle.numerics import F16, from_bits, rn, to_bits

ROWS, KEYS, D = 128, 64, 128


def read_operand(buf, rows, ks, lbo, sbo):
    """UMMA canonical K-major, no swizzle: -> float32 [rows, ks] from a uint8 buffer."""
    r = np.arange(rows)[:, None]
    k = np.arange(ks)[None, :]
    addr = (k // 8) * lbo + (r // 8) * sbo + (r % 8) * 16 + (k % 8) * 2
    bits = buf[addr].astype(np.uint16) | (buf[addr + 1].astype(np.uint16) << 8)
    return from_bits(bits, F16)


def put16(buf, addr, halves):
    """st.shared.v4.b32: 8 T values (as float32 exact in T) at byte address addr."""
    b = to_bits(np.asarray(halves, np.float32), F16).astype(np.uint16)
    buf[addr:addr + 16:2] = b & 0xFF
    buf[addr + 1:addr + 16:2] = b >> 8


def pf_load8(pool_row, kv_bits, dim0, scale, zero):
    """8 consecutive dims of one token, de-quantised: fma_T(T(q), scale, zero)."""
    if kv_bits == 16:
        raw = pool_row[dim0 * 2:dim0 * 2 + 16]
        return from_bits(raw[0::2].astype(np.uint16) | (raw[1::2].astype(np.uint16) << 8), F16)
    if kv_bits == 8:
        q = pool_row[dim0:dim0 + 8].astype(np.float32)
    else:
        w = int(pool_row[dim0 // 2:dim0 // 2 + 4].view(np.uint32)[0])
        n = [(w >> (4 * j)) & 15 for j in range(8)]
        # cvt_u4x8: registers (n0,n4) (n1,n5) (n2,n6) (n3,n7) -> natural dim order 0..7 by the KV nibble order
        q = np.array([n[0], n[4], n[1], n[5], n[2], n[6], n[3], n[7]], np.float32)
    return OA.t_fma(q, np.float32(scale), np.float32(zero), F16)


def emulate(pool, block_bytes, tables, cu_q, cu_k, layer_offset, block_len, Hq, Hkv, kv_bits, q_rot, qk_scale, v_pad=0):
    v_sbo = 128 + v_pad                  # kPfVSbo / kPfVLbo of the kernel (kPfVPad = v_pad)
    v_lbo = (D // 8) * v_sbo
    G = Hq // Hkv
    TQ = ROWS // G
    kRow = D * kv_bits // 8
    head_data_sz, head_param_sz = block_len * kRow, block_len * 4
    batch = len(cu_q) - 1
    T = cu_q[-1]
    out = np.zeros((T, Hq * D), np.float32)
    grid_x = T // TQ + batch
    for h in range(Hkv):
        for bx in range(grid_x):
            x, b = bx, 0
            while b < batch:
                nt = (cu_q[b + 1] - cu_q[b] + TQ - 1) // TQ
                if x < nt:
                    break
                x -= nt
                b += 1
            if b >= batch:
                continue
            q0 = x * TQ
            q_len, k_len = cu_q[b + 1] - cu_q[b], cu_k[b + 1] - cu_k[b]
            hist = k_len - q_len
            nq = min(TQ, q_len - q0)
            last_pos = hist + q0 + nq - 1
            n_tiles = last_pos // KEYS + 1
            r = np.arange(ROWS)
            tq, g = r // G, r % G
            valid = tq < nq
            pos_q = np.where(valid, hist + q0 + tq, 0x7fffffff)
            tok = cu_q[b] + q0 + tq
            q_op = np.zeros(ROWS * D * 2, np.uint8)
            for row in range(ROWS):
                for c in range(D // 8):
                    v = q_rot[tok[row], h * G + g[row], c * 8:c * 8 + 8] if valid[row] else np.zeros(8, np.float32)
                    put16(q_op, (c * ROWS + row) * 16, v)
            Q = read_operand(q_op, ROWS, D, ROWS * 16, 128).astype(np.float64)
            o = np.zeros((ROWS, D), np.float64)
            m = np.full(ROWS, -np.inf)
            l = np.zeros(ROWS)
            for j in range(n_tiles):
                key0 = j * KEYS
                blk = tables[b][key0 // block_len] * block_bytes + layer_offset
                bt0 = key0 % block_len
                kd = blk + (h * 2) * head_data_sz + bt0 * kRow
                vd = kd + head_data_sz
                kp = blk + Hkv * 2 * head_data_sz + (h * 2) * head_param_sz + bt0 * 4
                vp = kp + head_param_sz
                k_op = np.zeros(KEYS * D * 2, np.uint8)
                v_op = np.zeros((KEYS // 8) * v_lbo, np.uint8)

                def params(base, key):
                    if kv_bits == 16:
                        return 0.0, 0.0
                    pr = from_bits(pool[base + key * 4: base + key * 4 + 4].view(np.uint16), F16)
                    return pr[0], pr[1]

                for tid in range(128):
                    key, half = tid & 63, tid >> 6
                    live = key0 + key <= last_pos
                    s, z = params(kp, key) if live else (0.0, 0.0)
                    for cc in range(8):
                        x8 = pf_load8(pool[kd + key * kRow: kd + (key + 1) * kRow], kv_bits, half * 64 + cc * 8, s, z) \
                            if live else np.zeros(8, np.float32)
                        put16(k_op, ((half * 8 + cc) * KEYS + key) * 16, x8)
                for tid in range(128):
                    kg, dg = tid >> 4, tid & 15
                    a = np.zeros((8, 8), np.float32)          # a[jj][dim]: registers a[jj][i>>1] hold dims (2(i>>1), +1)
                    for jj in range(8):
                        key = kg * 8 + jj
                        if key0 + key <= last_pos:
                            s, z = params(vp, key)
                            a[jj] = pf_load8(pool[vd + key * kRow: vd + (key + 1) * kRow], kv_bits, dg * 8, s, z)
                    for i in range(8):
                        # y[q] = byte_perm(a[2q][i>>1], a[2q+1][i>>1], sel): (key 2q, key 2q+1) of dim i
                        y = [a[kk][(i >> 1) * 2 + (i & 1)] for kk in range(8)]
                        put16(v_op, kg * v_lbo + dg * v_sbo + i * 16, y)
                K = read_operand(k_op, KEYS, D, KEYS * 16, 128).astype(np.float64)       # B of S: N = keys, K = d
                Vt = read_operand(v_op, D, KEYS, v_lbo, v_sbo).astype(np.float64)       # B of PV: N = d, K = keys
                S = (Q @ K.T).astype(np.float32)
                kidx = key0 + np.arange(KEYS)[None, :]
                Sm = np.where(kidx <= pos_q[:, None], S, -np.inf)
                m_new = np.maximum(m, Sm.max(1))
                alpha = np.exp2((m - m_new) * qk_scale)
                P = np.exp2(Sm.astype(np.float64) * qk_scale - (m_new * qk_scale)[:, None])
                l = l * alpha + P.sum(1)
                m = m_new
                Pt = rn(P.astype(np.float32), F16).astype(np.float64)
                o = o * alpha[:, None] + Pt @ Vt.T
            res = rn((o / l[:, None]).astype(np.float32), F16)
            for row in range(ROWS):
                if valid[row]:
                    hq = h * G + g[row]
                    out[tok[row], hq * D:(hq + 1) * D] = res[row]
    return out


@pytest.mark.parametrize("kv_bits,heads,v_pad", [(8, (4, 1), 0), (4, (4, 1), 0), (16, (4, 1), 0), (8, (2, 2), 0),
                                                 (4, (6, 1), 0), (16, (6, 1), 0), (8, (4, 1), 16), (16, (2, 2), 16)])
def test_kernel_index_math_reproduces_the_oracle(kv_bits, heads, v_pad):
    Hq, Hkv = heads
    rng = np.random.default_rng(kv_bits * 10 + Hq)
    q_lens, hists = np.array([37, 1, 70]), np.array([0, 70, 30])
    k_lens = q_lens + hists
    block_len, layers, layer = 64, 2, 1
    layout = OA.KvLayout(Hkv, block_len, D, kv_bits)
    blocks_per = [(int(k) + block_len - 1) // block_len for k in k_lens]
    nblk = sum(blocks_per) + 1
    cache = OA.PagedKvCache(nblk, layers, layout, F16)
    cache.pool[:] = rng.integers(0, 256, cache.pool.shape, dtype=np.uint8)       # garbage, incl. NaN bit patterns
    perm = rng.permutation(nblk)
    tables, o = [], 0
    for n in blocks_per:
        tables.append(perm[o:o + n].tolist())
        o += n
    rope = OA.RopeParam(type=OA.ROPE_DEFAULT, dim=D, base=1e4)
    for b, hlen in enumerate(hists):
        if hlen:
            OA.process_kv(cache, tables[b], layer, 0, rn(rng.standard_normal((hlen, Hkv, D)).astype(np.float32), F16),
                          rn(rng.standard_normal((hlen, Hkv, D)).astype(np.float32), F16), rope)
    T = int(q_lens.sum())
    qkv = rn(rng.standard_normal((T, (Hq + 2 * Hkv) * D)).astype(np.float32), F16)
    ref = OA.prefill_attention(qkv, cache, tables, q_lens, k_lens, layer, Hq, Hkv, rope)      # also writes the chunk's K/V
    # what the prologue kernel hands the attention kernel: rotated Q (the pages were just written by the oracle)
    q_rot = np.zeros((T, Hq, D), np.float32)
    t0 = 0
    for b, ql in enumerate(q_lens):
        pos = (hists[b] + np.arange(ql))[:, None]
        q_rot[t0:t0 + ql] = OA.apply_rope(qkv[t0:t0 + ql, :Hq * D].reshape(ql, Hq, D), np.broadcast_to(pos, (ql, Hq)), rope, F16)
        t0 += ql
    cu_q = np.concatenate([[0], np.cumsum(q_lens)]).tolist()
    cu_k = np.concatenate([[0], np.cumsum(k_lens)]).tolist()
    c = float(np.float32(np.float32(1.4426950408889634) / np.float32(np.sqrt(np.float32(D)))))
    got = emulate(cache.pool.reshape(-1), cache.block_bytes, tables, cu_q, cu_k, cache.layer_offset(layer), block_len, Hq, Hkv,
                  kv_bits, q_rot, c, v_pad)
    assert np.isfinite(got).all()
    # tiled online softmax vs the oracle's single pass: a few fp16 ulps of an O(1) output
    assert np.abs(got - ref).max() <= 2e-3
