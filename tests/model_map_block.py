"""NumPy float32 model of the arithmetic of tpf::map_lin2_kernel (commpy_b200/csrc/bcjr.cu): four combined branch weights
per step relative to the most likely (u, cp), metrics rescaled every 4th step with a per-step redo of a block whose sum
decays below 2^-20, a-posteriori sums with the prior folded in and their 2^80 rescue.  Whole-frame recursions (no windows):
the model is about dynamic range and the rescaling logic, which the window split does not touch.  `per_step=True` gives the
arithmetic of the previous kernel's scaling (every step).  Used by tests/test_model_map.py (CPU) to show the kernel's
arithmetic meets the stated LLR tolerance against the fp64 oracle, including on frames that force the fallback."""
import numpy as np

F = np.float32
LOG2E, LN2 = F(1.4426950408889634), F(0.6931471805599453)
FLOOR = F(2.0 ** -20)


def _weights(ys, yp, la, c2):
    l = np.clip(la * LOG2E, F(-60), F(60)).astype(F)
    lu = (ys * c2 + l).astype(F)                        # (no clamp on the channel terms: 2^-|x| of a huge |x| is 0)
    lb = (yp * c2).astype(F)
    eu, eb = np.exp2(-np.abs(lu)).astype(F), np.exp2(-np.abs(lb)).astype(F)
    one = np.ones_like(eu)
    u1, u0 = np.where(lu >= 0, one, eu), np.where(lu >= 0, eu, one)
    q1, q0 = np.where(lb >= 0, one, eb), np.where(lb >= 0, eb, one)
    return np.stack([u0 * q0, u0 * q1, u1 * q0, u1 * q1], -1).astype(F), l      # (..., 4), index u << 1 | cp


def map_decode_model(ys, yp, trellis, noise_variance, La, per_step=False):
    """(batch, N) float arrays -> full a-posteriori LLR (batch, N) float32."""
    ys, yp, La = (np.asarray(a, dtype=F) for a in (ys, yp, La))
    nxt, out = np.asarray(trellis.next_state_table), np.asarray(trellis.output_table)
    S = nxt.shape[0]
    assert all((out[s, u] >> 1) == u for s in range(S) for u in range(2)), "systematic trellis"
    batch, N = ys.shape
    assert N % 4 == 0
    c2 = F(2.0) * (LOG2E / F(noise_variance))
    g, l = _weights(ys, yp, La, c2)                       # (batch, N, 4)
    with np.errstate(under="ignore"):
        def beta_raw(B, gw):
            Bn = np.empty_like(B)
            for s in range(S):
                Bn[:, s] = gw[:, out[s, 0]] * B[:, nxt[s, 0]] + gw[:, out[s, 1]] * B[:, nxt[s, 1]]
            return Bn

        def alpha_raw(A, gw):
            tx = np.empty((batch, 2 * S), F)
            for s in range(S):
                for u in range(2):
                    tx[:, 2 * s + u] = A[:, s] * gw[:, out[s, u]]
            An = np.zeros_like(A)
            for s in range(S):
                for u in range(2):
                    An[:, nxt[s, u]] += tx[:, 2 * s + u]
            return An, tx

        def scale(v):
            return (v * (F(1) / np.maximum(v.sum(1, dtype=F), F(1e-37)))[:, None]).astype(F)

        beta = np.empty((N + 1, batch, S), F)
        B = np.full((batch, S), F(1.0 / S))
        fb_b = 0
        for e1 in range(N, 0, -4):
            B0 = B
            for t in range(e1, e1 - 4, -1):
                beta[t] = B
                B = beta_raw(B, g[:, t - 1])
                if per_step:
                    B = scale(B)
            if not per_step:
                bad = ~(B.sum(1, dtype=F) >= FLOOR)
                if bad.any():
                    fb_b += int(bad.sum())
                    Bs = B0[bad]
                    for t in range(e1, e1 - 4, -1):
                        beta[t][bad] = Bs
                        Bs = scale(beta_raw_sub(Bs, g[bad, t - 1], nxt, out))
                    B = B.copy(); B[bad] = Bs
                B = scale(B)
        A = np.zeros((batch, S), F); A[:, 0] = 1
        D = np.empty((batch, N), F)
        fb_a = 0

        def app(tx, bt):
            a0 = np.zeros(tx.shape[0], F); a1 = np.zeros(tx.shape[0], F)
            for s in range(S):
                a0 += tx[:, 2 * s] * bt[:, nxt[s, 0]]
                a1 += tx[:, 2 * s + 1] * bt[:, nxt[s, 1]]
            small = np.maximum(a0, a1) < F(2.0 ** -40)
            if small.any():
                big = F(2.0 ** 40)
                b0 = np.zeros(tx.shape[0], F); b1 = np.zeros(tx.shape[0], F)
                for s in range(S):
                    b0 += (tx[:, 2 * s] * big) * (bt[:, nxt[s, 0]] * big)
                    b1 += (tx[:, 2 * s + 1] * big) * (bt[:, nxt[s, 1]] * big)
                a0, a1 = np.where(small, b0, a0), np.where(small, b1, a1)
            return (np.log2(np.maximum(a1, F(1e-37))) - np.log2(np.maximum(a0, F(1e-37)))).astype(F)

        for e0 in range(0, N, 4):
            A0 = A
            for t in range(e0 + 1, e0 + 5):
                A, tx = alpha_raw(A, g[:, t - 1])
                D[:, t - 1] = app(tx, beta[t])
                if per_step:
                    A = scale(A)
            if not per_step:
                bad = ~(A.sum(1, dtype=F) >= FLOOR)
                if bad.any():
                    fb_a += int(bad.sum())
                    As = A0[bad]
                    for t in range(e0 + 1, e0 + 5):
                        An, tx = alpha_raw_sub(As, g[bad, t - 1], nxt, out)
                        D[bad, t - 1] = app(tx, beta[t][bad])
                        As = scale(An)
                    A = A.copy(); A[bad] = As
                A = scale(A)
    L = (La + LN2 * (D - l)).astype(F)
    map_decode_model.fallbacks = (fb_b, fb_a)
    return L


def beta_raw_sub(B, gw, nxt, out):
    S = nxt.shape[0]
    Bn = np.empty_like(B)
    for s in range(S):
        Bn[:, s] = gw[:, out[s, 0]] * B[:, nxt[s, 0]] + gw[:, out[s, 1]] * B[:, nxt[s, 1]]
    return Bn


def alpha_raw_sub(A, gw, nxt, out):
    S = nxt.shape[0]
    tx = np.empty((A.shape[0], 2 * S), F)
    for s in range(S):
        for u in range(2):
            tx[:, 2 * s + u] = A[:, s] * gw[:, out[s, u]]
    An = np.zeros_like(A)
    for s in range(S):
        for u in range(2):
            An[:, nxt[s, u]] += tx[:, 2 * s + u]
    return An, tx
