"""Live pin of the oracle: where the reference checkout is present (the build container; never the GPU box), fresh random
cases -- not the committed goldens -- are run through the UNMODIFIED reference and through oracle/commpy_oracle.c and must
agree (bit-exact for Viterbi and LDPC min-sum incl. out_llrs, 1e-9 for the float outputs).  Skipped elsewhere."""
import os
import time
import warnings

import numpy as np
import pytest

from oracle import oracle, refimport

pytestmark = [pytest.mark.skipif(not refimport.available(), reason="reference checkout not present"),
              pytest.mark.filterwarnings("ignore")]


@pytest.fixture(scope="module")
def ref():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        refimport.import_reference()
        import commpy.channelcoding as rcc
        import commpy.modulation as rmod
    return rcc, rmod


def _seed():
    return int(time.time()) % (1 << 31)          # a different draw every run; printed on failure


def test_viterbi_and_bcjr_fresh_cases(ref):
    rcc, _ = ref
    seed = _seed()
    rs = np.random.RandomState(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        k7 = rcc.Trellis(np.array([6]), np.array([[0o133, 0o171]]))
        rsc = rcc.Trellis(np.array([3]), np.array([[1, 0o15]]), np.array([[0o13]]), "rsc")
    for mode in ("hard", "soft", "unquantized"):
        for term, tb in (("cont", None), ("term", 9)):
            msg = rs.randint(0, 2, 48)
            c = rcc.conv_encode(msg, k7, term).astype(float)
            if mode == "hard":
                x = np.abs(c - (rs.rand(len(c)) < 0.1))
            elif mode == "soft":
                x = (2 * c - 1) * 2 + rs.randn(len(c)) * 2.5
            else:
                x = (2 * c - 1) + rs.randn(len(c))
            want = rcc.viterbi_decode(x.copy(), k7, tb, mode)
            assert np.array_equal(want, oracle.viterbi_decode(x.copy(), k7, tb, mode)), (seed, mode, term, tb)
    N = 40
    s2 = 0.9
    sys_, par, la = (-1 + np.sqrt(s2) * rs.randn(N) for _ in range(3))
    with np.errstate(all="ignore"):
        want = rcc.map_decode(sys_, par, rsc, s2, 0.2 * la, "decode")
    got = oracle.map_decode(sys_, par, rsc, s2, 0.2 * la, "decode")
    assert np.allclose(want[0], got[0], rtol=1e-9, atol=1e-9) and np.array_equal(want[1], got[1]), seed


def test_ldpc_and_demapper_fresh_cases(ref):
    rcc, rmod = ref
    seed = _seed() + 1
    rs = np.random.RandomState(seed)
    design = os.path.join(refimport.REF_ROOT, "commpy/channelcoding/designs/ldpc/gallager/96.33.964.txt")
    params = rcc.get_ldpc_code_params(design)
    sigma = 0.9
    llr = 2.0 * (1.0 + sigma * rs.randn(2 * 96)) / sigma ** 2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want_dec, want_out = rcc.ldpc_bp_decode(llr.copy(), params, "MSA", 6)
    dec, out = oracle.ldpc_bp_decode(llr.copy(), params, "MSA", 6)
    assert np.array_equal(want_dec, dec) and np.array_equal(want_out, out), seed
    modem = rmod.QAMModem(16)
    y = modem.modulate(rs.randint(0, 2, 4 * 30)) + 0.4 * (rs.randn(30) + 1j * rs.randn(30))
    with np.errstate(all="ignore"):
        want = modem.demodulate(y, "soft", 0.7)
    got = oracle.demodulate(modem, y, "soft", 0.7)
    fin = np.isfinite(want)
    assert np.allclose(want[fin], got[fin], rtol=1e-9, atol=1e-9), seed
    assert np.array_equal(modem.demodulate(y, "hard"), oracle.demodulate(modem, y, "hard")), seed


def test_mimo_mirrors_fresh_cases(ref):
    """host-side MIMO mirrors (kbest, best_first_detector, idd_decoder) against the live reference on fresh draws"""
    import importlib
    _, rmod = ref
    rlinks = importlib.import_module("commpy.links")
    from commpy_b200 import modulation as mod
    from commpy_b200.links import idd_decoder
    seed = _seed()
    rs = np.random.RandomState(seed)
    for m, n in ((4, 2), (16, 3), (16, 4)):
        rq = rmod.QAMModem(m)
        mq = mod.QAMModem(m)
        np.testing.assert_array_equal(rq.constellation, mq.constellation)

        def demode(s):
            return rq.demodulate(s, "hard")
        for _ in range(10):
            nb = rq.num_bits_symbol
            h = (rs.randn(n, n) + 1j * rs.randn(n, n)) * np.sqrt(0.5)
            nv = n / 10 ** (rs.uniform(5, 25) / 10)
            y = h.dot(rq.modulate(rs.randint(0, 2, n * nb))) + (rs.randn(n) + 1j * rs.randn(n)) * np.sqrt(nv / 2)
            msg = "seed %d" % seed
            np.testing.assert_allclose(mod.kbest(y, h, mq.constellation, 6), rmod.kbest(y, h, rq.constellation, 6), atol=1e-12,
                                       err_msg=msg)
            np.testing.assert_allclose(mod.kbest(y, h, mq.constellation, 6, nv, "soft", demode),
                                       rmod.kbest(y, h, rq.constellation, 6, nv, "soft", demode), rtol=1e-9, atol=1e-9, err_msg=msg)
            stack = tuple(rs.randint(1, 6, n - 1))
            np.testing.assert_allclose(mod.best_first_detector(y, h, mq.constellation, stack, nv, demode, 100),
                                       rmod.best_first_detector(y, h, rq.constellation, stack, nv, demode, 100), rtol=1e-9,
                                       atol=1e-9, err_msg=msg)
    # idd_decoder: same closures through both, 3 iterations
    rq = rmod.QAMModem(16)

    def detector(y, h, constellation, noise_var, a_priori):
        return rmod.kbest(y, h, constellation, 8, noise_var, "soft", lambda s: rq.demodulate(s, "hard")) + 0.25 * a_priori

    def decoder(llr):
        return np.tanh(llr) * 3 + np.roll(llr, 1)

    def decision(llr):
        return (llr > 0).astype(int)
    nvec, n = 5, 2
    h = (rs.randn(nvec, n, n) + 1j * rs.randn(nvec, n, n)) * np.sqrt(0.5)
    y = np.einsum("ijk,ik->ij", h, rq.modulate(rs.randint(0, 2, nvec * n * 4)).reshape(nvec, n)) + 0.1 * rs.randn(nvec, n)
    ap = rs.randn(nvec * n * 4)
    got = idd_decoder(detector, decoder, decision, 3)(y, h, rq.constellation, 0.02, ap.copy(), n * 4)
    want = rlinks.idd_decoder(detector, decoder, decision, 3)(y, h, rq.constellation, 0.02, ap.copy(), n * 4)
    np.testing.assert_array_equal(got, want)
