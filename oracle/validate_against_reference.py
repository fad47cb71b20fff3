"""Pin the CPU oracle against the UNMODIFIED reference and write the golden fixtures.

Runs only in the build container (needs /root/reference).  For every hot-path function it feeds seeded
inputs to the imported reference (veeresht/CommPy @ 9aecd7c) and to oracle/commpy_oracle.c, asserts
equality (bit-exact for integer outputs, 1e-9 relative for float64 outputs) and stores inputs + reference
outputs in tests/golden/*.npz.  tests/test_oracle_golden.py replays those fixtures everywhere (no reference
needed); the GPU parity tests use them too.

    python oracle/validate_against_reference.py            # ~3-4 min, rewrites tests/golden/
"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]
sys.path.insert(0, ROOT)
from oracle import oracle, refimport  # noqa: E402

warnings.simplefilter("ignore")
GOLD = os.path.join(ROOT, "tests", "golden")


def trellis_specs():
    """(name, memory, g_matrix, feedback, code_type, polynomial_format)"""
    a = np.array
    return [
        ("t57", a([2]), a([[5, 7]]), None, "default", "MSB"),
        ("rsc_legacy", a([2]), a([[1, 7]]), 5, "rsc", "MSB"),
        ("r23", a([2, 1]), a([[5, 7, 0], [0, 2, 3]]), None, "default", "MSB"),
        ("r23_lsb", a([2, 1]), a([[5, 7, 0], [0, 2, 6]]), None, "default", "LSB"),
        ("r23_rsc", a([1, 1]), a([[1, 0, 0], [0, 1, 3]]), a([[2, 2], [3, 1]]), "rsc", "MSB"),
        ("k7", a([6]), a([[0o133, 0o171]]), None, "default", "MSB"),
        ("k7_wifi_quirk", a([6]), a([[133, 171]]), None, "default", "MSB"),
        ("rsc_k4", a([3]), a([[1, 0o15]]), a([[0o13]]), "rsc", "MSB"),
    ]


def make_trellis(mod, spec):
    _, mem, g, fb, ct, pf = spec
    fb = fb.copy() if isinstance(fb, np.ndarray) else fb
    return mod.Trellis(mem.copy(), g.copy(), fb, ct, pf)


def main():
    t_start = time.time()
    commpy = refimport.import_reference()
    import commpy.channelcoding as rcc
    import commpy.modulation as rmod
    os.makedirs(GOLD, exist_ok=True)

    # ---------------- Trellis tables + conv_encode KATs ----------------
    tabs = {}
    for spec in trellis_specs():
        tr = make_trellis(rcc, spec)
        tabs[spec[0] + "_next"] = tr.next_state_table
        tabs[spec[0] + "_out"] = tr.output_table
        msg = np.array((0, 0, 1, 0, 1, 1, 0, 1, 0, 0, 1, 1)[: 12 - 12 % tr.k])
        tabs[spec[0] + "_enc_cont"] = rcc.conv_encode(msg, tr, "cont")
        tabs[spec[0] + "_enc_term"] = rcc.conv_encode(msg, tr, "term")
        tabs[spec[0] + "_msg"] = msg
    np.savez_compressed(os.path.join(GOLD, "trellis_tables.npz"), **tabs)
    print("trellis tables: %d specs" % len(trellis_specs()))

    # ---------------- Viterbi ----------------
    rs = np.random.RandomState(20240901)
    vit = {}
    case = 0
    for spec in trellis_specs():
        tr = make_trellis(rcc, spec)
        big = spec[0] in ("k7", "k7_wifi_quirk")
        for mode in ("hard", "soft", "unquantized"):
            for term in ("cont", "term"):
                for tb in ((None, 15, 7) if big else (None, 15)):
                    nbits = (200 if big else 96) - (200 if big else 96) % tr.k
                    if big and mode == "hard" and tb is None and term == "cont":
                        nbits = 1024                                     # the C1 shape
                    msg = rs.randint(0, 2, nbits)
                    c = rcc.conv_encode(msg, tr, term).astype(float)
                    if mode == "hard":
                        x = np.abs(c - (rs.rand(len(c)) < (0.12 if case % 2 else 0.03)))
                    elif mode == "soft":
                        x = (2 * c - 1) * 2 + rs.randn(len(c)) * 2.5
                    else:
                        x = (2 * c - 1) + rs.randn(len(c)) * 1.0
                    ref = rcc.viterbi_decode(x.copy(), tr, tb, mode)
                    orc = oracle.viterbi_decode(x.copy(), tr, tb, mode)
                    assert np.array_equal(ref, orc), ("viterbi", spec[0], mode, term, tb)
                    key = "c%03d" % case
                    vit[key + "_x"] = x
                    vit[key + "_ref"] = ref.astype(np.int8)
                    vit[key + "_meta"] = np.array([spec[0], mode, term, str(tb)])
                    case += 1
    # +-inf LLRs (test_convcode.py:168-178)
    tr = make_trellis(rcc, trellis_specs()[5])
    msg = rs.randint(0, 2, 120)
    c = rcc.conv_encode(msg, tr, "cont").astype(float)
    x = np.where(c == 1.0, np.inf, -np.inf)
    ref = rcc.viterbi_decode(x.copy(), tr, 15, "soft")
    assert np.array_equal(ref, oracle.viterbi_decode(x.copy(), tr, 15, "soft")) and np.array_equal(ref, msg)
    vit["c%03d_x" % case] = x
    vit["c%03d_ref" % case] = ref.astype(np.int8)
    vit["c%03d_meta" % case] = np.array(["k7", "soft", "cont", "15"])
    case += 1
    np.savez_compressed(os.path.join(GOLD, "viterbi.npz"), **vit)
    print("viterbi: %d cases bit-exact (%.0fs)" % (case, time.time() - t_start))

    # ---------------- BCJR / turbo ----------------
    bc = {}
    case = 0
    for name in ("rsc_k4", "rsc_legacy", "t57"):
        spec = [s for s in trellis_specs() if s[0] == name][0]
        tr = make_trellis(rcc, spec)
        for N, eb in ((64, 0.0), (256, 1.0), (512, 2.0)):
            msg = rs.randint(0, 2, N)
            coded = rcc.conv_encode(msg, tr, "cont")
            sigma2 = 1.0 / (2 * 0.5 * 10 ** (eb / 10))
            sys_ = 2.0 * coded[0::2] - 1 + np.sqrt(sigma2) * rs.randn(N)
            par = 2.0 * coded[1::2] - 1 + np.sqrt(sigma2) * rs.randn(N)
            La = rs.randn(N) * (1.5 if case % 2 else 0.0)
            for mode in ("decode", "compute"):
                Lr, br = rcc.map_decode(sys_, par, tr, sigma2, La, mode)
                Lo, bo = oracle.map_decode(sys_, par, tr, sigma2, La, mode)
                assert np.array_equal(br, bo), ("map bits", name, N)
                assert np.allclose(Lr, Lo, rtol=1e-9, atol=1e-9), ("map L", name, N, np.abs(Lr - Lo).max())
            key = "m%02d" % case
            bc[key + "_sys"], bc[key + "_par"], bc[key + "_La"] = sys_, par, La
            bc[key + "_L"], bc[key + "_bits"] = Lr if mode == "decode" else Lr, br
            Lr, br = rcc.map_decode(sys_, par, tr, sigma2, La, "decode")
            bc[key + "_L"], bc[key + "_bits"] = Lr, br.astype(np.int8)
            bc[key + "_meta"] = np.array([name, str(sigma2)])
            case += 1
    print("map_decode: %d cases (%.0fs)" % (case, time.time() - t_start))
    tcase = 0
    spec = [s for s in trellis_specs() if s[0] == "rsc_k4"][0]
    tr = make_trellis(rcc, spec)
    for N, eb, iters in ((128, 0.5, 3), (256, 1.0, 6), (512, 1.5, 4)):
        il = rcc.RandInterlv(N, 1)
        msg = rs.randint(0, 2, N)
        s_, p1, p2 = rcc.turbo_encode(msg, tr, tr, il)
        s_, p1, p2 = s_[:N], p1[:N], p2[:N]
        sigma2 = 1.0 / (2 * (1 / 3) * 10 ** (eb / 10))
        ys = 2.0 * s_ - 1 + np.sqrt(sigma2) * rs.randn(N)
        y1 = 2.0 * p1 - 1 + np.sqrt(sigma2) * rs.randn(N)
        y2 = 2.0 * p2 - 1 + np.sqrt(sigma2) * rs.randn(N)
        ref = rcc.turbo_decode(ys, y1, y2, tr, sigma2, iters, il)
        orc = oracle.turbo_decode(ys, y1, y2, tr, sigma2, iters, il)
        assert np.array_equal(ref, orc), ("turbo", N)
        key = "t%02d" % tcase
        bc[key + "_sys"], bc[key + "_p1"], bc[key + "_p2"] = ys, y1, y2
        bc[key + "_perm"] = il.p_array.astype(np.int32)
        bc[key + "_bits"] = ref.astype(np.int8)
        bc[key + "_msg"] = msg.astype(np.int8)
        bc[key + "_meta"] = np.array([str(sigma2), str(iters)])
        tcase += 1
    np.savez_compressed(os.path.join(GOLD, "bcjr_turbo.npz"), **bc)
    print("turbo_decode: %d cases bit-exact (%.0fs)" % (tcase, time.time() - t_start))

    # ---------------- LDPC min-sum ----------------
    ld = {}
    case = 0
    ddir = os.path.join(refimport.REF_ROOT, "commpy", "channelcoding", "designs", "ldpc")
    for rel, nblk, iters, eb in (("gallager/96.33.964.txt", 1, 8, 1.0), ("gallager/96.33.964.txt", 3, 25, 2.5),
                                 ("gallager/96.3.963.txt", 2, 10, 2.0), ("wimax/1440.720.txt", 1, 6, 1.0),
                                 ("wimax/960.720.a.txt", 2, 5, 2.5)):
        params = rcc.get_ldpc_code_params(os.path.join(ddir, rel), compute_matrix=True)
        n = params["n_vnodes"]
        rate = 1.0 - params["n_cnodes"] / n
        sigma = 1.0 / np.sqrt(2 * rate * 10 ** (eb / 10))
        llr = 2.0 * (1.0 + sigma * rs.randn(n * nblk)) / sigma ** 2
        if case == 1:
            llr[:5] = (1000.0, -1000.0, 0.0, -0.0, 3.0)          # clipping and zeros
        a, b = llr.copy(), llr.copy()
        dr, lr = rcc.ldpc_bp_decode(a, params, "MSA", iters)
        do, lo = oracle.ldpc_bp_decode(b, params, "MSA", iters)
        assert np.array_equal(dr, do), ("ldpc dec", rel)
        assert np.array_equal(lr, lo), ("ldpc llr", rel, np.abs(lr - lo).max())
        assert np.array_equal(a, b)                              # in-place clip
        H = params["parity_check_matrix"].tocsr()
        H.sort_indices()
        key = "l%02d" % case
        ld[key + "_llr"] = llr
        ld[key + "_dec"] = dr
        ld[key + "_out"] = lr
        ld[key + "_indptr"] = H.indptr.astype(np.int32)
        ld[key + "_indices"] = H.indices.astype(np.int32)
        ld[key + "_meta"] = np.array([rel, str(nblk), str(iters), str(H.shape[0]), str(H.shape[1])])
        case += 1
    # sum-product variant (ldpc.py:209-227): decisions equal, out_llrs to 1e-9 (the reference goes through complex log2/exp2)
    scase = 0
    for rel, nblk, iters, eb in (("gallager/96.33.964.txt", 2, 20, 2.5), ("wimax/1440.720.txt", 1, 8, 1.5),
                                 ("wimax/960.720.a.txt", 1, 6, 3.0)):
        params = rcc.get_ldpc_code_params(os.path.join(ddir, rel), compute_matrix=True)
        n = params["n_vnodes"]
        rate = 1.0 - params["n_cnodes"] / n
        sigma = 1.0 / np.sqrt(2 * rate * 10 ** (eb / 10))
        llr = 2.0 * (1.0 + sigma * rs.randn(n * nblk)) / sigma ** 2
        a, b = llr.copy(), llr.copy()
        with np.errstate(all="ignore"):
            dr, lr = rcc.ldpc_bp_decode(a, params, "SPA", iters)
        do, lo = oracle.ldpc_bp_decode(b, params, "SPA", iters)
        assert np.array_equal(dr, do), ("ldpc spa dec", rel)
        assert np.allclose(lr, lo, rtol=1e-9, atol=1e-9), ("ldpc spa llr", rel)
        H = params["parity_check_matrix"].tocsr()
        H.sort_indices()
        key = "s%02d" % scase
        ld[key + "_llr"], ld[key + "_dec"], ld[key + "_out"] = llr, dr, lr
        ld[key + "_indptr"], ld[key + "_indices"] = H.indptr.astype(np.int32), H.indices.astype(np.int32)
        ld[key + "_meta"] = np.array([rel, str(nblk), str(iters), str(H.shape[0]), str(H.shape[1])])
        scase += 1
    np.savez_compressed(os.path.join(GOLD, "ldpc.npz"), **ld)
    print("ldpc_bp_decode MSA: %d cases exact incl. out_llrs, SPA: %d cases (%.0fs)" % (case, scase, time.time() - t_start))

    # ---------------- demapper ----------------
    dm = {}
    case = 0
    from itertools import product
    custom = [re + im * 1j for re, im in product((-3.5, -0.5, 0.5, 3.5), repeat=2)]       # test_modulation.py:95
    modems = [("psk4", rmod.PSKModem(4)), ("psk8", rmod.PSKModem(8)), ("psk16", rmod.PSKModem(16)),
              ("qam4", rmod.QAMModem(4)), ("qam16", rmod.QAMModem(16)), ("qam64", rmod.QAMModem(64)),
              ("qam256", rmod.QAMModem(256)), ("custom16", rmod.Modem(custom))]
    for name, md in modems:
        dm[name + "_constellation"] = np.asarray(md.constellation, dtype=np.complex128)
        nsym = 24 if md.m >= 64 else 40
        for nv in (0.1, 1.0, 4.0):
            bits = rs.randint(0, 2, nsym * md.num_bits_symbol)
            y = md.modulate(bits) + np.sqrt(nv * md.Es / 10) * (rs.randn(nsym) + 1j * rs.randn(nsym))
            with np.errstate(all="ignore"):
                lr = md.demodulate(y, "soft", nv)
            lo = oracle.demodulate(md, y, "soft", nv)
            fin = np.isfinite(lr)
            assert np.allclose(lr[fin], lo[fin], rtol=1e-9, atol=1e-9), ("demod soft", name, nv)
            hr = md.demodulate(y, "hard")
            ho = oracle.demodulate(md, y, "hard")
            assert np.array_equal(hr, ho), ("demod hard", name)
            key = "d%02d" % case
            dm[key + "_y"], dm[key + "_llr"], dm[key + "_hard"] = y, lr, hr.astype(np.int8)
            dm[key + "_meta"] = np.array([name, str(nv)])
            case += 1
    np.savez_compressed(os.path.join(GOLD, "demod.npz"), **dm)
    print("demodulate: %d cases (%.0fs)" % (case, time.time() - t_start))
    print("oracle pinned against the reference; fixtures in", GOLD)


if __name__ == "__main__":
    main()
