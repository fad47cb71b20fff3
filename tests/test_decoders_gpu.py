"""GPU parity for BCJR/turbo, LDPC min-sum, the demapper and the error counter: CUDA kernels (through the
CommPy-shaped wrappers) against the golden vectors of the reference and against the CPU oracle."""
import os
from itertools import product

import numpy as np
import pytest
import torch

import helpers
from oracle import oracle
from commpy_b200.channelcoding import (RandInterlv, conv_encode, ldpc_bp_decode, ldpc_bp_decode_batch, map_decode,
                                        map_decode_batch, turbo_decode, turbo_decode_batch, turbo_encode)
from commpy_b200.modulation import Modem, PSKModem, QAMModem

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# stated tolerances
MAP_ATOL, MAP_RTOL = 1e-4, 1e-4
DEMAP_ATOL, DEMAP_RTOL = 5e-4, 5e-4


def _spec_trellis(name):
    from test_host_mirror import _specs
    import warnings
    from commpy_b200.channelcoding import Trellis
    mem, g, fb, ct, pf = _specs()[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return Trellis(mem, g, fb, ct, pf)


# ---------------------------------------------------------------- BCJR / turbo
def test_map_decode_golden_llr_tolerance():
    g = np.load(os.path.join(GOLD, "bcjr_turbo.npz"))
    worst = 0.0
    for c in range(9):
        name, s2 = g["m%02d_meta" % c]
        tr = _spec_trellis(str(name))
        L, bits = map_decode(g["m%02d_sys" % c], g["m%02d_par" % c], tr, float(s2), g["m%02d_La" % c], "decode")
        ref = g["m%02d_L" % c]
        fin = np.isfinite(ref)
        err = np.abs(L[fin] - ref[fin]) - MAP_RTOL * np.abs(ref[fin])
        worst = max(worst, float(err.max()))
        assert (err <= MAP_ATOL).all(), (name, float(err.max()))
        safe = np.abs(ref) > 1e-3
        assert np.array_equal(bits[safe], g["m%02d_bits" % c][safe])
        L2, b2 = map_decode(g["m%02d_sys" % c], g["m%02d_par" % c], tr, float(s2), g["m%02d_La" % c], "compute")
        assert np.allclose(L2, L) and not b2.any() and b2.dtype == np.dtype("int")
    print("map_decode worst excess error", worst)


def test_map_decode_batch_vs_oracle_long_frames():
    tr = helpers.rsc_k4()
    rs = np.random.RandomState(21)
    N, batch = 6144, 6
    msgs = rs.randint(0, 2, (batch, N))
    coded = np.stack([conv_encode(m, tr, "cont") for m in msgs])
    s2 = 1.0 / (2 * 0.5 * 10 ** (1.0 / 10))
    ys = 2.0 * coded[:, 0::2] - 1 + np.sqrt(s2) * rs.randn(batch, N)
    yp = 2.0 * coded[:, 1::2] - 1 + np.sqrt(s2) * rs.randn(batch, N)
    La = rs.randn(batch, N)
    L, bits = map_decode_batch(ys, yp, tr, s2, La, "decode")
    L = L.cpu().numpy()
    for b in range(batch):
        Lo, bo = oracle.map_decode(ys[b], yp[b], tr, s2, La[b], "decode")
        assert (np.abs(L[b] - Lo) <= MAP_ATOL + MAP_RTOL * np.abs(Lo)).all(), float(np.abs(L[b] - Lo).max())


def test_map_block_rescaling_matches_per_step_and_oracle_on_contradicted_frames():
    """The default MAP kernel rescales its metrics every 4th step and redoes a block step by step when the metric sum
    decays (bcjr.cu, map_lin2_kernel).  Frames whose observations contradict every trellis path (random signs instead of
    code words, strong random priors) force that fallback -- tests/test_model_map.py counts it on the same cases with the
    NumPy model of the kernel's arithmetic.  Both kernel forms must agree with the fp64 oracle within the MAP tolerance."""
    from commpy_b200 import _lib
    tr = helpers.rsc_k4()
    rs = np.random.RandomState(77)
    N, batch = 2048, 40
    for s2, amp, lamp in helpers.CONTRADICTED_MAP_CASES:
        ys = amp * rs.choice([-1.0, 1.0], (batch, N)) + np.sqrt(s2) * rs.randn(batch, N)
        yp = amp * rs.choice([-1.0, 1.0], (batch, N)) + np.sqrt(s2) * rs.randn(batch, N)
        La = lamp * rs.randn(batch, N)
        res = {}
        for per_step in (1, 0):
            _lib.set_option(_lib.OPT_BCJR_PER_STEP_SCALING, per_step)
            try:
                L, bits = map_decode_batch(ys, yp, tr, s2, La, "decode")
            finally:
                _lib.set_option(_lib.OPT_BCJR_PER_STEP_SCALING, 0)
            res[per_step] = (L.cpu().numpy().astype(np.float64), bits.cpu().numpy())
        assert np.isfinite(res[0][0]).all()
        for b in range(0, batch, 8):
            Lo, bo = oracle.map_decode(ys[b], yp[b], tr, s2, La[b], "decode")
            ok = np.isfinite(Lo) & (np.abs(Lo) < 40.0)
            assert ok.mean() > 0.9
            for k in (0, 1):
                d = np.abs(res[k][0][b][ok] - Lo[ok])
                assert (d <= MAP_ATOL + MAP_RTOL * np.abs(Lo[ok])).all(), (k, s2, float(d.max()))
                big = ok & (np.abs(Lo) > 1e-2)
                assert np.array_equal(res[k][1][b][big], bo[big])


def test_turbo_decode_golden_and_batch():
    g = np.load(os.path.join(GOLD, "bcjr_turbo.npz"))
    tr = helpers.rsc_k4()

    class IL:
        pass
    tot = bad = 0
    for c in range(3):
        s2, iters = g["t%02d_meta" % c]
        il = IL()
        il.p_array = g["t%02d_perm" % c]
        bits = turbo_decode(g["t%02d_sys" % c], g["t%02d_p1" % c], g["t%02d_p2" % c], tr, float(s2), int(iters), il)
        assert bits.dtype == np.dtype("int")
        tot += bits.size
        bad += int((bits != g["t%02d_bits" % c]).sum())
    assert bad / tot <= 1e-3, (bad, tot)
    # batch at the C3 frame length: agreement with the oracle and with the message
    rs = np.random.RandomState(22)
    N, batch = 6144, 8
    il = RandInterlv(N, 1)
    s2 = 1.0 / (2 * (1 / 3) * 10 ** (1.0 / 10))
    ys, y1, y2, msgs = [], [], [], []
    for b in range(batch):
        msg = rs.randint(0, 2, N)
        s_, p1, p2 = turbo_encode(msg, tr, tr, il)
        ys.append(2.0 * s_[:N] - 1 + np.sqrt(s2) * rs.randn(N))
        y1.append(2.0 * p1[:N] - 1 + np.sqrt(s2) * rs.randn(N))
        y2.append(2.0 * p2[:N] - 1 + np.sqrt(s2) * rs.randn(N))
        msgs.append(msg)
    ys, y1, y2, msgs = map(np.stack, (ys, y1, y2, msgs))
    got = turbo_decode_batch(ys, y1, y2, tr, s2, 6, il).cpu().numpy()
    want = oracle.turbo_decode_batch(ys, y1, y2, tr, s2, 6, il, threads=8)
    assert (got == want).mean() >= 0.999, float((got == want).mean())
    assert abs(int((got != msgs).sum()) - int((want != msgs).sum())) <= max(10, 0.05 * (want != msgs).sum())


def test_turbo_step_major_loop_shapes_priors_and_trellises():
    """The turbo loop on step-major arrays (the default where the block-rescaled kernel applies) against the frame-major loop
    with its interleaver kernels (CPB_OPT_TURBO_FRAME_MAJOR: same MAP arithmetic, so the SAME bits) and against the fp64
    oracle: ragged batches (frames padded to 32), frame lengths with 4-step tails, a single group, an a-priori L_int, 4- and
    8-state RSC codes, short MAP windows."""
    from commpy_b200 import _lib
    from commpy_b200.channelcoding import Trellis, set_map_window
    rs = np.random.RandomState(31)
    k4 = helpers.rsc_k4()
    k3 = Trellis(np.array([2]), np.array([[1, 5]]), np.array([[7]]), "rsc")
    cases = [(k4, 5, 100, 3, False, 0), (k4, 33, 516, 4, True, 0), (k4, 70, 2048, 6, False, 0), (k4, 3, 4, 2, True, 0),
             (k3, 37, 1000, 5, True, 0), (k4, 40, 6144, 6, False, 128), (k4, 9, 3072, 4, True, 256)]
    for tr, batch, N, iters, prior, window in cases:
        il = RandInterlv(N, 3)
        s2 = 1.0 / (2 * (1 / 3) * 10 ** (1.5 / 10))
        ys, y1, y2, msgs = [], [], [], []
        for b in range(batch):
            msg = rs.randint(0, 2, N)
            s_, p1, p2 = turbo_encode(msg, tr, tr, il)
            ys.append(2.0 * s_[:N] - 1 + np.sqrt(s2) * rs.randn(N))
            y1.append(2.0 * p1[:N] - 1 + np.sqrt(s2) * rs.randn(N))
            y2.append(2.0 * p2[:N] - 1 + np.sqrt(s2) * rs.randn(N))
            msgs.append(msg)
        ys, y1, y2, msgs = map(np.stack, (ys, y1, y2, msgs))
        La = (1.5 * (2.0 * msgs - 1) + rs.randn(batch, N)) if prior else None      # a helpful, noisy prior
        got = {}
        set_map_window(window)
        try:
            for fm in (1, 0):
                _lib.set_option(_lib.OPT_TURBO_FRAME_MAJOR, fm)
                got[fm] = turbo_decode_batch(ys, y1, y2, tr, s2, iters, il, La).cpu().numpy()
        finally:
            _lib.set_option(_lib.OPT_TURBO_FRAME_MAJOR, 0)
            set_map_window(0)
        assert np.array_equal(got[0], got[1]), (batch, N, int((got[0] != got[1]).sum()))
        want = np.stack([oracle.turbo_decode(ys[b], y1[b], y2[b], tr, s2, iters, il, None if La is None else La[b])
                         for b in range(min(batch, 6))])
        agree = (got[0][:len(want)] == want).mean()
        assert agree >= 0.999 or N < 64, (batch, N, float(agree))
        if N < 64:
            assert (got[0][:len(want)] != want).sum() <= 1


def test_release_scratch_between_calls():
    """cpb_release_scratch hands the library's scratch pool back to the driver; the next call allocates again and decodes the same."""
    from commpy_b200 import _lib
    tr = helpers.rsc_k4()
    rs = np.random.RandomState(5)
    ys, y1, y2 = (rs.randn(40, 1024) - 1 for _ in range(3))
    il = RandInterlv(1024, 7)
    a = turbo_decode_batch(ys, y1, y2, tr, 0.7, 3, il).cpu().numpy()
    torch.cuda.synchronize()
    _lib.release_scratch()
    b = turbo_decode_batch(ys, y1, y2, tr, 0.7, 3, il).cpu().numpy()
    assert np.array_equal(a, b)


# ---------------------------------------------------------------- LDPC
def _golden_ldpc(c):
    import scipy.sparse as sp
    g = np.load(os.path.join(GOLD, "ldpc.npz"))
    rel, nblk, iters, m, n = g["l%02d_meta" % c]
    H = sp.csr_matrix((np.ones(len(g["l%02d_indices" % c]), np.int8), g["l%02d_indices" % c], g["l%02d_indptr" % c]),
                      shape=(int(m), int(n)))
    return g, {"n_vnodes": int(n), "n_cnodes": int(m), "parity_check_matrix": H.tocsc()}, int(iters)


def test_ldpc_fp64_bit_exact_with_reference():
    for c in range(5):
        g, params, iters = _golden_ldpc(c)
        llr = g["l%02d_llr" % c].copy()
        dec, out = ldpc_bp_decode(llr, params, "MSA", iters)          # default precision: fp64 parity mode
        assert dec.dtype == np.int8
        assert np.array_equal(dec, g["l%02d_dec" % c]), c
        assert np.array_equal(out, g["l%02d_out" % c]), (c, float(np.abs(out - g["l%02d_out" % c]).max()))
        assert np.abs(llr).max() <= 500.0                               # clipped in place (ldpc.py:186)
    with pytest.raises(NameError):
        ldpc_bp_decode(np.zeros(params["n_vnodes"]), params, "XYZ", 3)


def test_ldpc_fp32_converged_frames_match():
    g, params, _ = _golden_ldpc(3)                    # WiMax 1440.720
    n = params["n_vnodes"]
    rs = np.random.RandomState(23)
    sigma = 1.0 / np.sqrt(2 * 0.5 * 10 ** (2.5 / 10))
    batch = 48
    llr = (2.0 * (1.0 + sigma * rs.randn(batch, n)) / sigma ** 2)
    want_dec, want_out, want_it = oracle.ldpc_bp_decode(llr.reshape(-1).copy(), params, "MSA", 50, return_iters=True,
                                                        threads=8)
    dec, out, it = ldpc_bp_decode_batch(llr.astype(np.float32), params, 50, "fp32", return_iters=True)
    dec, out, it = dec.cpu().numpy(), out.cpu().numpy(), it.cpu().numpy()
    conv = want_it < 50
    assert conv.sum() >= batch // 2
    want_dec = want_dec.reshape(n, batch).T if batch > 1 else want_dec[None]
    want_out = want_out.reshape(n, batch).T if batch > 1 else want_out[None]
    assert np.array_equal(dec[conv], want_dec[conv])
    assert np.array_equal(it[conv], want_it[conv])
    assert np.allclose(out[conv], want_out[conv], rtol=1e-4, atol=1e-3)
    # fp64 mode: every frame, converged or not, exact
    dec64, out64, it64 = ldpc_bp_decode_batch(llr, params, 50, "fp64", return_iters=True)
    assert np.array_equal(dec64.cpu().numpy(), want_dec)
    assert np.array_equal(out64.cpu().numpy(), want_out)
    assert np.array_equal(it64.cpu().numpy(), want_it)


@pytest.mark.parametrize("batch", [256, 384])
def test_ldpc_bulk_copy_check_pass_exact(batch, monkeypatch):
    """Batches of >= 128 frames run the check pass staged by the bulk-copy engine (cn_bulk_kernel, chunk-major R):
    fp64 stays bit-exact with the reference on every frame (converged or not), and fp32 output is bit-identical to the
    register-staged kernels (same operations in the same order)."""
    g, params, _ = _golden_ldpc(3)                    # WiMax 1440.720
    n = params["n_vnodes"]
    rs = np.random.RandomState(99 + batch)
    sigma = 1.0 / np.sqrt(2 * 0.5 * 10 ** (1.6 / 10))
    llr = (2.0 * (1.0 + sigma * rs.randn(batch, n)) / sigma ** 2)
    llr[5] = 2.0 * (1.0 + 3.0 * rs.randn(n))          # a frame that never converges
    want_dec, want_out, want_it = oracle.ldpc_bp_decode(llr.reshape(-1).copy(), params, "MSA", 20, return_iters=True,
                                                        threads=8)
    want_dec = want_dec.reshape(n, batch).T
    want_out = want_out.reshape(n, batch).T
    assert (want_it < 20).sum() > batch // 4 and (want_it == 20).sum() >= 1
    from commpy_b200 import _lib
    _lib.set_option(_lib.OPT_LDPC_NO_BULK, 0)
    dec64, out64, it64 = ldpc_bp_decode_batch(llr.copy(), params, 20, "fp64", return_iters=True)
    assert np.array_equal(dec64.cpu().numpy(), want_dec)
    assert np.array_equal(out64.cpu().numpy(), want_out)
    assert np.array_equal(it64.cpu().numpy(), want_it)
    dec32, out32, it32 = ldpc_bp_decode_batch(llr.astype(np.float32), params, 20, "fp32", return_iters=True)
    _lib.set_option(_lib.OPT_LDPC_NO_BULK, 1)
    try:
        dec32r, out32r, it32r = ldpc_bp_decode_batch(llr.astype(np.float32), params, 20, "fp32", return_iters=True)
    finally:
        _lib.set_option(_lib.OPT_LDPC_NO_BULK, 0)
    assert torch.equal(dec32, dec32r) and torch.equal(it32, it32r)
    assert torch.equal(out32.view(torch.int32), out32r.view(torch.int32))


def test_ldpc_noiseless_roundtrip_and_zero_iterations():
    """test_ldpc.py:77-106: a noiseless codeword needs zero iterations and out_llrs stays the clipped input."""
    g, params, _ = _golden_ldpc(3)
    n = params["n_vnodes"]
    llr = np.full(n, 1000.0)
    dec, out = ldpc_bp_decode(llr, params, "MSA", 10)
    assert not dec.any() and np.array_equal(out, np.full(n, 500.0)) and llr.max() == 500.0


# ---------------------------------------------------------------- demapper
def test_demod_soft_golden_tolerance_and_hard_exact():
    g = np.load(os.path.join(GOLD, "demod.npz"))
    custom = [re + im * 1j for re, im in product((-3.5, -0.5, 0.5, 3.5), repeat=2)]
    mods = {"psk4": PSKModem(4), "psk8": PSKModem(8), "psk16": PSKModem(16), "qam4": QAMModem(4),
            "qam16": QAMModem(16), "qam64": QAMModem(64), "qam256": QAMModem(256), "custom16": Modem(custom)}
    worst = 0.0
    for c in range(24):
        name, nv = g["d%02d_meta" % c]
        md = mods[str(name)]
        y = g["d%02d_y" % c]
        ref = g["d%02d_llr" % c]
        out = md.demodulate(y, "soft", float(nv))
        assert out.dtype == np.float64 and out.shape == ref.shape
        fin = np.isfinite(ref)
        err = np.abs(out[fin] - ref[fin]) - DEMAP_RTOL * np.abs(ref[fin])
        worst = max(worst, float(err.max()))
        assert (err <= DEMAP_ATOL).all(), (name, nv, float(err.max()))
        assert np.isfinite(out).all()
        hard = md.demodulate(y, "hard")
        assert hard.dtype == np.int8 and np.array_equal(hard, g["d%02d_hard" % c]), name
    print("demod worst excess error", worst)
    with pytest.raises(ValueError):
        mods["qam16"].demodulate(np.zeros(4, complex), "medium")


def test_demod_roundtrip_all_patterns():
    """test_modulation.py:159-162: modulate -> hard demodulate is the identity for every bit pattern."""
    for md in (QAMModem(4), QAMModem(16), QAMModem(64), PSKModem(4), PSKModem(16), PSKModem(64)):
        for bits in product(*((0, 1),) * md.num_bits_symbol):
            assert np.array_equal(bits, md.demodulate(md.modulate(bits), "hard"))


def test_demod_separable_equals_general_path():
    import torch
    q = QAMModem(256)
    gen = Modem(q.constellation, reorder_as_gray=False)
    gen._constellation = gen._constellation * (1 + 0j)
    rs = np.random.RandomState(24)
    y = (rs.randn(4096) + 1j * rs.randn(4096)) * 9
    a = q.demodulate(y, "soft", 12.0)
    # force the general kernel by perturbing one point by one ulp-ish amount
    pts = np.array(q.constellation, dtype=np.complex128)
    pts[3] += 1e-6
    gen.constellation = pts
    b = gen.demodulate(y, "soft", 12.0)
    assert np.allclose(a, b, rtol=2e-3, atol=2e-3)


# ---------------------------------------------------------------- error counter
def test_count_errors():
    import ctypes as C
    import torch
    from commpy_b200 import _lib
    rs = np.random.RandomState(25)
    a = rs.randint(0, 2, (513, 1030)).astype(np.uint8)
    b = a.copy()
    flips = rs.rand(*a.shape) < 0.01
    flips[::7] = False
    b[flips] ^= 1
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    rc = _lib.load().cpb_count_errors(_lib.ptr(ta), _lib.ptr(tb), C.c_int64(513), C.c_int64(1030), C.c_int64(1030),
                                      C.c_int64(1030), _lib.ptr(cnt), _lib.stream_ptr(torch))
    _lib.check(rc, "count")
    got = cnt.cpu().numpy()
    assert got[0] == int(flips.sum()) and got[1] == int(flips.any(axis=1).sum())


def test_ldpc_dvbs2_size_surrogate_exact_fp64_and_converged_fp32():
    """C4 shape: (64800, 32400), 226,799 edges (DVB-S2-SHAPED surrogate, helpers.dvbs2_like_H)."""
    H = helpers.dvbs2_like_H()
    params = {"n_vnodes": 64800, "n_cnodes": 32400, "parity_check_matrix": H.tocsc()}
    rs = np.random.RandomState(26)
    frames = []
    for eb in (1.0, 2.6, 2.6, 3.0):          # one frame that does not converge, three that do
        sigma = 1.0 / np.sqrt(2 * 0.5 * 10 ** (eb / 10))
        frames.append(2.0 * (1.0 + sigma * rs.randn(64800)) / sigma ** 2)
    llr = np.stack(frames)
    want_dec, want_out, want_it = oracle.ldpc_bp_decode(llr.reshape(-1).copy(), params, "MSA", 12, return_iters=True,
                                                        threads=4)
    want_dec, want_out = want_dec.T, want_out.T
    dec, out, it = ldpc_bp_decode_batch(llr.copy(), params, 12, "fp64", return_iters=True)
    assert np.array_equal(it.cpu().numpy(), want_it)
    assert np.array_equal(dec.cpu().numpy(), want_dec)
    assert np.array_equal(out.cpu().numpy(), want_out)
    dec32, out32, it32 = ldpc_bp_decode_batch(llr.astype(np.float32), params, 12, "fp32", return_iters=True)
    conv = want_it < 12
    assert conv.sum() >= 2
    assert np.array_equal(dec32.cpu().numpy()[conv], want_dec[conv])
    assert np.array_equal(it32.cpu().numpy()[conv], want_it[conv])


def test_map_decode_generic_trellis_fallback_kernel():
    """A trellis that is not one of the compile-time instances goes through the table-driven lane-per-state kernel."""
    from commpy_b200.channelcoding import Trellis
    rs = np.random.RandomState(27)
    for tr in (Trellis(np.array([3]), np.array([[0o15, 0o17]])), Trellis(np.array([4]), np.array([[0o23, 0o35]]))):
        N = 400
        msg = rs.randint(0, 2, N)
        coded = conv_encode(msg, tr, "cont")
        s2 = 0.7
        ys = 2.0 * coded[0::2] - 1 + np.sqrt(s2) * rs.randn(N)
        yp = 2.0 * coded[1::2] - 1 + np.sqrt(s2) * rs.randn(N)
        La = rs.randn(N)
        L, bits = map_decode(ys, yp, tr, s2, La, "decode")
        Lo, bo = oracle.map_decode(ys, yp, tr, s2, La, "decode")
        assert (np.abs(L - Lo) <= MAP_ATOL + MAP_RTOL * np.abs(Lo)).all(), float(np.abs(L - Lo).max())
        safe = np.abs(Lo) > 1e-3
        assert np.array_equal(bits[safe], bo[safe])


def test_map_decode_unaligned_length_scalar_path():
    tr = helpers.rsc_k4()
    rs = np.random.RandomState(28)
    for N in (37, 1537, 2049):                   # not multiples of 4; 1537 and 2049 are split into windows
        msg = rs.randint(0, 2, N)
        coded = conv_encode(msg, tr, "cont")
        s2 = 0.8
        ys = 2.0 * coded[0::2] - 1 + np.sqrt(s2) * rs.randn(N)
        yp = 2.0 * coded[1::2] - 1 + np.sqrt(s2) * rs.randn(N)
        La = 0.5 * rs.randn(N)
        L, _ = map_decode(ys, yp, tr, s2, La, "decode")
        Lo, _ = oracle.map_decode(ys, yp, tr, s2, La, "decode")
        assert (np.abs(L - Lo) <= MAP_ATOL + MAP_RTOL * np.abs(Lo)).all(), (N, float(np.abs(L - Lo).max()))


def test_ldpc_spa_matches_reference_golden():
    """'SPA' (ldpc.py:209-227) in fp64: same decisions, out_llrs to 1e-6 (formula identical, rounding differs)."""
    import scipy.sparse as sp
    g = np.load(os.path.join(GOLD, "ldpc.npz"))
    for c in range(3):
        rel, nblk, iters, m, n = g["s%02d_meta" % c]
        H = sp.csr_matrix((np.ones(len(g["s%02d_indices" % c]), np.int8), g["s%02d_indices" % c], g["s%02d_indptr" % c]),
                          shape=(int(m), int(n)))
        params = {"n_vnodes": int(n), "n_cnodes": int(m), "parity_check_matrix": H.tocsc()}
        dec, out = ldpc_bp_decode(g["s%02d_llr" % c].copy(), params, "SPA", int(iters))
        assert np.array_equal(dec, g["s%02d_dec" % c]), rel
        assert np.allclose(out, g["s%02d_out" % c], rtol=1e-6, atol=1e-6), (rel, float(np.abs(out - g["s%02d_out" % c]).max()))
    # fp32 throughput mode: decisions of converged blocks agree
    dec32, _ = ldpc_bp_decode(g["s00_llr"].copy(), _golden_ldpc(0)[1], "SPA", 20, precision="fp32")
    assert (dec32 == g["s00_dec"]).mean() > 0.99


def test_edge_cases_other_decoders():
    import torch
    # demapper: zero and one symbol
    q = QAMModem(16)
    assert q.demodulate(np.zeros(0, complex), "soft", 1.0).shape == (0,)
    one = q.demodulate(np.array([0.3 - 1.2j]), "soft", 0.7)
    assert np.allclose(one, oracle.demodulate(q, np.array([0.3 - 1.2j]), "soft", 0.7), rtol=5e-4, atol=5e-4)
    # LDPC: zero iterations returns the (clipped) channel decisions, single block, batch that is not a multiple of 4
    g, params, _ = _golden_ldpc(0)
    llr = g["l00_llr"].copy()
    dec, out = ldpc_bp_decode(llr.copy(), params, "MSA", 0)
    assert np.array_equal(dec, np.signbit(llr).astype(np.int8)) and np.array_equal(out, np.clip(llr, -500, 500))
    rs = np.random.RandomState(29)
    x = rs.randn(7, 96) * 3
    want_dec, want_out = oracle.ldpc_bp_decode(x.reshape(-1).copy(), params, "MSA", 15)
    dec, out = ldpc_bp_decode_batch(x.copy(), params, 15, "fp64")
    assert np.array_equal(dec.cpu().numpy(), want_dec.T) and np.array_equal(out.cpu().numpy(), want_out.T)
    # MAP: a single very short frame, and a batch of 3
    tr = helpers.rsc_k4()
    for N, batch in ((4, 1), (12, 3)):
        ys, yp, La = rs.randn(batch, N), rs.randn(batch, N), rs.randn(batch, N)
        L, bits = map_decode_batch(ys, yp, tr, 0.9, La)
        for b in range(batch):
            Lo, bo = oracle.map_decode(ys[b], yp[b], tr, 0.9, La[b])
            assert np.allclose(L[b].cpu().numpy(), Lo, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("algorithm,precision", [("MSA", "fp32"), ("MSA", "fp64"), ("SPA", "fp32")])
def test_ldpc_fer_acceptance_of_the_reference(algorithm, precision):
    """commpy/channelcoding/tests/test_ldpc.py:28-66 on the GPU path: Gallager (96, 48) code, all-zero codeword over BPSK-AWGN,
    100 iterations, frame error rate ~0.2 at Eb/N0 = 2.0 dB and ~0.1 at 2.5 dB (the reference's own tolerance: rtol 0.6).
    The reference estimates each FER from 50 frame errors; here every point is 8,192 frames."""
    import torch
    g, params, _ = _golden_ldpc(0)                    # Gallager 96.33.964
    n = params["n_vnodes"]
    assert n == 96
    rs = np.random.RandomState(321)
    rate, Es = 0.5, 1.0
    fer = []
    for ebno in (2.0, 2.5):
        noise_std = 1 / np.sqrt((10 ** (ebno / 10.0)) * rate * 2 / Es)
        rx = 1.0 + noise_std * rs.randn(8192, n)
        llr = 2.0 * rx / noise_std ** 2
        dec = ldpc_bp_decode_batch(llr.astype(np.float64 if precision == "fp64" else np.float32), params, 100, precision,
                                   return_llrs=False, decoder_algorithm=algorithm)
        fer.append(float((dec.sum(dim=1) > 0).float().mean()))
    np.testing.assert_allclose(fer, (0.2, 0.1), rtol=0.6, atol=0, err_msg=algorithm + " does not perform as expected")
    assert fer[1] < fer[0]
