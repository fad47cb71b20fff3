"""NumPy model of the history-key Viterbi kernel (commpy_b200/csrc/viterbi.cu, namespace jump).

Checked on CPU against the oracle before any GPU time is spent.  The ALGORITHM (64-state k=1, n=2 feed-forward
code, CommPy Trellis convention):

  key    = metric << FB | field            FB = 6 + B bits of path history (B = 4 trellis steps per block)
  field  bit i  <->  input u_{t0-5+i}  on the survivor path of the state, t0 = last block boundary
         (bits 0..5 are the survivor's state at t0, bit 6+j is the input of sub-step j of the block)
  ACS    new state ns = l + 32u:  min( K[2l] + Bm[out(2l,u)],  K[2l+1] + Bm[out(2l+1,u)] ) + (u << (6+j))
         One add + one add-min per state and step, nothing else: the smaller KEY wins, and on equal metrics the
         fields decide -- the two candidates agree on the newer history bits and differ in the predecessor's
         LSB, so predecessor 2l wins: the reference's first-minimum rule (convcode.py:612-642).
  best   min over the 64 keys: lowest metric, ties -> lowest field = lowest state index (np.argmin, :645); the
         minimum key's field is the whole path of the best state back to the block boundary.
  block  every B steps: store the low B bits of every key (the survivor's inputs u_{t0-5}..u_{t0-2}: a B-step
         jump pointer), then key = metric | state.
  traceback  bit p is read on the path from best[min(p+D-1, T)] (App. A.1-8): start from the best key's field,
         then jump B steps per look-up:  state(t0-B) = ((state(t0) & 3) << 4) | nibble[t0][state(t0)].
"""
import numpy as np

B = 4
TBB = 24


def out_sym(M, G0, G1, s, u):
    v = u
    for b in range(1, M + 1):
        v |= ((s >> (M - b)) & 1) << b
    return ((bin(v & G0).count("1") & 1) << 1) | (bin(v & G1).count("1") & 1)


def acs_forward(coded, G0, G1, mode="hard", M=6, key_bits=16, qbits=None):
    """Forward pass.  Returns (bf, nib, T, L): bf[tau] = field of the best key after step tau (tau = 1..T),
    nib[b][s] = low B bits of state s's key at the end of block b (steps 4b+1 .. 4b+4)."""
    assert M == 6
    S, H, FB = 64, 32, 6 + B
    n_in = len(coded)
    L = n_in // 2
    T = L + M - 1
    otab = [[out_sym(M, G0, G1, s, u) for u in (0, 1)] for s in range(S)]
    if mode == "hard":
        big = 16
    else:
        raise NotImplementedError
    K = np.array([((0 if s == 0 else big) << FB) | s for s in range(S)], dtype=np.int64)
    bf = np.zeros(T + 1, dtype=np.int64)
    nblk = (T + B - 1) // B
    nib = np.zeros((nblk + 1, S), dtype=np.int64)
    lim = 1 << key_bits
    for tau in range(1, T + 1):
        j = (tau - 1) % B
        if tau <= L:
            r0, r1 = int(coded[2 * (tau - 1)]) & 1, int(coded[2 * (tau - 1) + 1]) & 1
        else:
            r0 = r1 = 0
        a = (r0 << 1) | r1
        Bm = [bin(o ^ a).count("1") << FB for o in range(4)]
        Kn = np.zeros(S, dtype=np.int64)
        for l in range(H):
            for u in (0, 1):
                c = K[2 * l + 1] + Bm[otab[2 * l + 1][u]] + (u << (6 + j))
                a_ = K[2 * l] + Bm[otab[2 * l][u]] + (u << (6 + j))
                assert c < lim and a_ < lim
                Kn[l + u * H] = min(a_, c)
        mn = int(Kn.min())
        bf[tau] = mn & ((1 << FB) - 1)
        if j == B - 1:
            blk = (tau - 1) // B
            nib[blk] = Kn & ((1 << B) - 1)
            fld = (Kn >> B) & 63
            assert np.array_equal(fld, np.arange(S))
            Kn = ((Kn >> FB) << FB) | np.arange(S)
            if tau % 16 == 0:
                Kn -= (mn >> FB) << FB
        K = Kn
    return bf, nib, T, L


def best_state(bf, tau):
    j = (tau - 1) % B
    return (int(bf[tau]) >> (j + 1)) & 63


def traceback_simple(bf, nib, T, L, D):
    """One independent jump traceback per output bit (the definition; slow)."""
    out = np.zeros(L, dtype=np.int64)
    for p in range(L):
        tstar = min(p + D - 1, T)
        j = (tstar - 1) % B
        t0 = tstar - 1 - j                      # block boundary below tstar
        reg = int(bf[tstar])                    # bit i <-> u_{t0-5+i}
        while t0 - 5 > p + 1:
            s = reg & 63
            blk = t0 // B - 1                   # the block that ended at step t0
            reg = (reg << B) | int(nib[blk][s])
            t0 -= B
        out[p] = (reg >> (p + 1 - (t0 - 5))) & 1
    return out


def traceback_blocks(bf, nib, T, L, D, stats=None, tbb=TBB):
    """The kernel's traceback: blocks of `tbb` windows, one shared walk per block (phase A) that retires the current
    path wherever it misses best[tau]; retired paths are finished with a fixed number of jumps (phase B)."""
    assert tbb % B == 0
    out = np.full(L, -1, dtype=np.int64)
    NJ = max(0, -(-(D - 8) // B))               # jumps that take any path retired at lo >= t0 below u_{lo-D+3}

    def jump(reg, t0):
        s = reg & 63
        return ((reg << B) | int(nib[t0 // B - 1][s])), t0 - B

    def finish(reg, t0, lo, hi, final_top):
        """path serving the windows (lo, hi]: bits u_{lo-D+3} .. u_{hi-D+2} (and everything up to u_L when it is
        the path from T)"""
        for _ in range(NJ):
            if t0 >= B:
                reg, t0 = jump(reg, t0)
        for tau in range(lo + 1, hi + 1):
            q = tau - D + 2                     # u_q = output bit q-1
            if q >= 1:
                out[q - 1] = (reg >> (q - (t0 - 5))) & 1
        if final_top:
            for q in range(hi - D + 3, L + 1):
                if q >= 1:
                    out[q - 1] = (reg >> (q - (t0 - 5))) & 1

    ts = D - 2
    ts -= ts % B                                # block-aligned start; windows below D-1 do not exist
    ts = max(ts, 0)
    while ts < T:
        te = min(ts + tbb, T)
        final = te == T
        hi = te
        j = (te - 1) % B
        t0 = te - 1 - j
        reg = int(bf[te])
        tasks = []
        first_of_path = True
        tau = te
        while tau > ts:
            j = (tau - 1) % B
            if tau - 1 - j != t0:               # crossed a block boundary: one jump
                reg, t0 = jump(reg, t0)
                assert t0 == tau - 1 - j
            if tau < hi:
                on_path = ((reg >> (j + 1)) & 63) == best_state(bf, tau)
                if not on_path:
                    tasks.append((reg, t0, tau, hi, final and hi == te))
                    hi = tau
                    reg = int(bf[tau])
            tau -= 1
        tasks.append((reg, t0, ts, hi, final and hi == te))
        if stats is not None:
            stats["blocks"] = stats.get("blocks", 0) + 1
            stats["tasks"] = stats.get("tasks", 0) + len(tasks) - 1
        for (r, t, lo, h, ft) in tasks:
            finish(r, t, lo, h, ft)
        ts = te
    # windows only exist for tau >= D-1: bits below are covered because ts started at or below D-2
    assert (out >= 0).all(), np.nonzero(out < 0)[0][:8]
    return out


def decode(coded, G0, G1, D=None, mode="hard", M=6, stats=None, simple=False):
    L = len(coded) // 2
    if D is None:
        D = min(5 * M, L)
    bf, nib, T, L = acs_forward(coded, G0, G1, mode, M)
    assert T >= D - 1
    if simple:
        return traceback_simple(bf, nib, T, L, D)
    return traceback_blocks(bf, nib, T, L, D, stats)
