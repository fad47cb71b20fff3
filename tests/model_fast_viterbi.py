"""NumPy model of viterbi_fast_kernel (commpy_b200/csrc/viterbi.cu): key-form ACS, finite sentinel,
survivor ring of D+9 steps, 16-window block traceback with per-window fallback, M-1 shortcut.
Used on CPU to check the kernel's ALGORITHM against the oracle before spending GPU time."""
import numpy as np

TBB = 16


def out_sym(M, G0, G1, s, u):
    v = u
    for b in range(1, M + 1):
        v |= ((s >> (M - b)) & 1) << b
    return ((bin(v & G0).count("1") & 1) << 1) | (bin(v & G1).count("1") & 1)


def decode(coded, G0, G1, D=None, mode="hard", qmax=1 << 19, M=6, stats=None):
    S = 1 << M
    H = S // 2
    n_in = len(coded)
    L = n_in // 2
    T = L + M - 1
    if D is None:
        D = min(5 * M, L)
    assert M + 1 <= D <= 48 and T >= D - 1
    R = TBB + D - 2 - (M - 1)
    ring_w = np.zeros((R, S), dtype=np.int64)
    ring_b = np.zeros(R, dtype=np.int64)
    out = np.full(L, -1, dtype=np.int64)
    if mode == "hard":
        big = 16 << 6
    else:
        x = np.asarray(coded, dtype=np.float32)
        if mode == "soft":
            x = np.clip(x, -500, 500)
        amax = max(float(np.max(np.abs(x))), 1e-30)
        scale = 2.0 ** np.floor(np.log2(qmax / amax))
        big = (2 * M * qmax + 1) << 6
        padq = -1.0 if mode == "unquantized" else 0.0
    K = np.array([(0 if s == 0 else big) | s for s in range(S)], dtype=np.int64)
    otab = [[out_sym(M, G0, G1, s, u) for u in (0, 1)] for s in range(S)]
    slot = 0
    next_te = D - 2 + TBB

    def tb_block(ts, te, final, slot_te):
        p0 = ts - D + 2
        tau_min = ts - D + 3 + (M - 1)
        q_hi = te if final else te - D + 2
        acc = {}
        cons = 0
        s = int(ring_b[slot_te])
        sl = slot_te
        for tau in range(te, tau_min - 1, -1):
            if tau > ts and s == int(ring_b[sl]):
                cons |= 1 << (tau - ts - 1)
            q = tau - (M - 1)
            if 1 <= q <= q_hi:
                acc[q - 1 - p0] = s & 1
            s = ((s << 1) & (S - 1)) | int(ring_w[sl, s])
            sl = R - 1 if sl == 0 else sl - 1
        nwin = te - ts
        for j in range(nwin):
            if (cons >> j) & 1:
                continue
            if stats is not None:
                stats["fallback"] = stats.get("fallback", 0) + 1
            tp = ts + 1 + j
            sl = (slot_te - (te - tp)) % R
            s2 = int(ring_b[sl])
            for _ in range(D - 2 - (M - 1)):
                s2 = ((s2 << 1) & (S - 1)) | int(ring_w[sl, s2])
                sl = R - 1 if sl == 0 else sl - 1
            acc[j] = s2 & 1
        cnt = (L - p0) if final else TBB
        for i in range(cnt):
            out[p0 + i] = acc[i]

    for tau in range(1, T + 1):
        if mode == "hard":
            if tau <= L:
                r0, r1 = int(coded[2 * (tau - 1)]) & 1, int(coded[2 * (tau - 1) + 1]) & 1
            else:
                r0 = r1 = 0
            a = (r0 << 1) | r1
            Bm = [bin(o ^ a).count("1") << 6 for o in range(4)]
        else:
            if tau <= L:
                r0, r1 = float(x[2 * (tau - 1)]), float(x[2 * (tau - 1) + 1])
            else:
                r0 = r1 = padq
            q0 = int(np.rint(np.clip(np.float32(r0) * np.float32(scale), -qmax, qmax)))
            q1 = int(np.rint(np.clip(np.float32(r1) * np.float32(scale), -qmax, qmax)))
            z0, o0, z1, o1 = max(q0, 0) << 6, max(-q0, 0) << 6, max(q1, 0) << 6, max(-q1, 0) << 6
            Bm = [z0 + z1, z0 + o1, o0 + z1, o0 + o1]
        Kn = np.zeros(S, dtype=np.int64)
        W = np.zeros(S, dtype=np.int64)
        for l in range(H):
            for u in (0, 1):
                m = min(K[2 * l] + Bm[otab[2 * l][u]], K[2 * l + 1] + Bm[otab[2 * l + 1][u]])
                ns = l + u * H
                Kn[ns] = (m & ~63) | ns
                W[ns] = m & 1
        mn = int(Kn.min())
        ring_w[slot] = W
        ring_b[slot] = mn & 63
        if tau % 16 == 0:
            Kn -= mn & ~63
        assert Kn.max() < (1 << (16 if mode == "hard" else 32))
        K = Kn
        if tau == T:
            tb_block(next_te - TBB, tau, True, slot)
        elif tau == next_te:
            tb_block(next_te - TBB, tau, False, slot)
            next_te += TBB
        slot = 0 if slot + 1 == R else slot + 1
    assert (out >= 0).all()
    return out
