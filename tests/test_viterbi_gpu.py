"""GPU parity: cpb_viterbi_decode (through the CommPy-shaped Python wrappers) against the CPU oracle."""
import numpy as np
import pytest

import helpers
from oracle import oracle
from commpy_b200.channelcoding import viterbi_decode, viterbi_decode_batch

pytestmark = pytest.mark.gpu


def _agree(got, want):
    return float((np.asarray(got) == np.asarray(want)).mean())


@pytest.mark.parametrize("make", [helpers.k7, helpers.k7_wifi_quirk])
@pytest.mark.parametrize("tb", [None, 15, 7, 46, 48])
@pytest.mark.parametrize("term", ["cont", "term"])
def test_k7_hard_bit_exact(make, tb, term):
    tr = make()
    rs = np.random.RandomState(11)
    for nbits, flip in ((200, 0.12), (1024, 0.03), (333, 0.06)):
        _, x = helpers.channel_frames(tr, rs, 70, nbits, "hard", term, flip=flip)
        want = oracle.viterbi_decode_batch(x, tr, tb, "hard")
        got = viterbi_decode_batch(x.astype(np.uint8), tr, tb, "hard")
        assert got.dtype == np.uint8 and got.shape == want.shape
        assert np.array_equal(got, want)


@pytest.mark.parametrize("mode", ["soft", "unquantized"])
@pytest.mark.parametrize("tb", [None, 15])
def test_k7_float_agreement(mode, tb):
    tr = helpers.k7()
    rs = np.random.RandomState(12)
    tot = 0
    bad = 0
    for nbits, eb in ((256, 0.0), (1024, 2.0), (1024, 4.0)):
        msgs, x = helpers.channel_frames(tr, rs, 64, nbits, mode, "cont", ebn0_db=eb)
        want = oracle.viterbi_decode_batch(x, tr, tb, mode)
        got = viterbi_decode_batch(x.astype(np.float32), tr, tb, mode)
        tot += want.size
        bad += int((got != want).sum())
        # BER against the transmitted message must match the oracle's within a few bits
        assert abs(int((got != msgs).sum()) - int((want != msgs).sum())) <= max(8, 0.02 * (want != msgs).sum())
    assert bad / tot <= 1e-4, "bit agreement %.6f" % (1 - bad / tot)


def test_generic_trellises_all_modes():
    rs = np.random.RandomState(13)
    for tr in helpers.reference_test_trellises() + [helpers.rsc_k4()]:
        for mode in ("hard", "soft", "unquantized"):
            for term in ("cont", "term"):
                nb = 120 * tr.k
                _, x = helpers.channel_frames(tr, rs, 33, nb, mode, term, flip=0.08, ebn0_db=1.0)
                for tb in (None, 15, 5):
                    want = oracle.viterbi_decode_batch(x, tr, tb, mode)
                    xin = x.astype(np.uint8) if mode == "hard" else x.astype(np.float32)
                    got = viterbi_decode_batch(xin, tr, tb, mode)
                    if mode == "hard":
                        assert np.array_equal(got, want), (tr.k, tr.n, mode, term, tb)
                    else:
                        assert _agree(got, want) >= 0.999, (tr.k, tr.n, mode, term, tb, _agree(got, want))


def test_single_frame_signature_and_errors():
    tr = helpers.k7()
    rs = np.random.RandomState(14)
    msgs, x = helpers.channel_frames(tr, rs, 1, 300, "hard", "term", flip=0.02)
    out = viterbi_decode(x[0], tr)
    assert out.ndim == 1 and out.dtype == np.dtype("int") and len(out) == 306
    assert np.array_equal(out, oracle.viterbi_decode(x[0], tr))
    with pytest.raises(ValueError):
        viterbi_decode(x[0], tr, decoding_type="bogus")
    with pytest.raises(ValueError):
        viterbi_decode(x[0] * 3, tr)
    with pytest.raises(ValueError):
        viterbi_decode(x[0][:20], tr, tb_depth=40)


def test_reference_roundtrips_inf_llr():
    """commpy/channelcoding/tests/test_convcode.py:133-178: noiseless / +-inf LLR round trips, tb_depth 15."""
    from commpy_b200.channelcoding import conv_encode
    rs = np.random.RandomState(17121996 % (2 ** 31))
    for tr in helpers.reference_test_trellises() + [helpers.k7()]:
        msg = rs.randint(0, 2, 1000 - 1000 % tr.k)
        coded = conv_encode(msg, tr)
        assert np.array_equal(viterbi_decode(coded.astype(float), tr, 15)[:len(msg)], msg)
        assert np.array_equal(viterbi_decode(2.0 * coded - 1, tr, 15, "unquantized")[:len(msg)], msg)
        assert np.array_equal(viterbi_decode((2.0 * coded - 1) * np.inf, tr, 15, "soft")[:len(msg)], msg)
        cont = conv_encode(msg, tr, termination="cont")
        noisy = 10.0 * cont - 5 + rs.randn(len(cont)) * 2
        assert np.array_equal(viterbi_decode(noisy, tr, 15, "soft"), msg)


def test_full_size_properties_and_kernel_cross_check():
    """BASELINE-size checks that need no oracle: (a) noiseless encode -> decode round trip over a full 65,536-frame
    batch, hard and soft; (b) the register-resident fast kernel and the table-driven generic kernel -- two independent
    implementations -- agree bit for bit on 8,192 noisy N=1024 frames (hard) and to 1e-4 (soft)."""
    import os
    import torch
    tr = helpers.k7()
    rs = np.random.RandomState(15)
    msgs = rs.randint(0, 2, (4096, 1024))
    coded = helpers.encode_batch(msgs, tr, "cont").astype(np.uint8)
    big = torch.from_numpy(coded).cuda().repeat(16, 1).contiguous()              # 65,536 frames
    out = viterbi_decode_batch(big, tr, None, "hard")
    want = torch.from_numpy(msgs.astype(np.uint8)).cuda().repeat(16, 1)
    # an unterminated ('cont') frame is padded with RECEIVED ZEROS in hard mode (convcode.py:727-728), which biases the
    # last few bits exactly as in the reference; everything before the last constraint length must be error free
    assert torch.equal(out[:, :1016], want[:, :1016])
    tail = oracle.viterbi_decode_batch(coded[:64].astype(np.float64), tr, None, "hard", threads=4)
    assert np.array_equal(out[:64].cpu().numpy(), tail)
    soft = (2.0 * big.float() - 1.0) * 4.0
    out = viterbi_decode_batch(soft, tr, None, "soft")
    assert torch.equal(out, want)
    del big, soft, out, want
    _, x = helpers.channel_frames(tr, rs, 8192, 1024, "hard", "cont", flip=0.06)
    xh = torch.from_numpy(x.astype(np.uint8)).cuda()
    fast = viterbi_decode_batch(xh, tr, None, "hard")
    _, xs = helpers.channel_frames(tr, rs, 4096, 1024, "soft", "cont", ebn0_db=2.0)
    xsf = torch.from_numpy(xs.astype(np.float32)).cuda()
    fast_s = viterbi_decode_batch(xsf, tr, None, "soft")
    from commpy_b200 import _lib
    _lib.set_option(_lib.OPT_VITERBI_FORCE_GENERIC, 1)
    try:
        gen = viterbi_decode_batch(xh, tr, None, "hard")
        gen_s = viterbi_decode_batch(xsf, tr, None, "soft")
    finally:
        _lib.set_option(_lib.OPT_VITERBI_FORCE_GENERIC, 0)
    assert torch.equal(fast, gen)
    assert (fast_s != gen_s).float().mean().item() <= 1e-4


def test_edge_cases_sizes_and_alignment():
    """Empty batch, single frame, batch sizes that do not fill a warp, odd input lengths (unaligned rows), the shortest
    frames / depths the reference semantics allow."""
    import torch
    tr = helpers.k7()
    rs = np.random.RandomState(16)
    # empty batch
    out = viterbi_decode_batch(np.zeros((0, 2048), np.uint8), tr, None, "hard")
    assert out.shape == (0, 1024)
    out = viterbi_decode_batch(torch.zeros((0, 64), dtype=torch.float32, device="cuda"), tr, None, "soft")
    assert tuple(out.shape) == (0, 32)
    for mode in ("hard", "soft", "unquantized"):
        for batch in (1, 31, 33, 65, 129):
            for nbits, term in ((40, "term"), (41, "cont"), (16, "cont"), (1023, "cont")):
                _, x = helpers.channel_frames(tr, rs, batch, nbits, mode, term, flip=0.05, ebn0_db=2.0)
                if nbits == 41:
                    x = x[:, :-1]                    # odd number of coded values: the last one is ignored (L = int(len/2))
                want = oracle.viterbi_decode_batch(x, tr, None, mode, threads=4)
                xin = x.astype(np.uint8) if mode == "hard" else x.astype(np.float32)
                got = viterbi_decode_batch(xin, tr, None, mode)
                if mode == "hard":
                    assert np.array_equal(got, want), (mode, batch, nbits)
                else:
                    assert (got == want).mean() >= 0.999, (mode, batch, nbits)
    # smallest depth (generic kernel: D = 2) and the largest the fast path takes (48), device-resident unaligned view
    _, x = helpers.channel_frames(tr, rs, 9, 200, "hard", "cont", flip=0.05)
    for tb in (2, 3, 6, 8, 9, 10, 11, 12, 13, 46, 47, 48, 49, 120):
        assert np.array_equal(viterbi_decode_batch(x.astype(np.uint8), tr, tb, "hard"),
                              oracle.viterbi_decode_batch(x, tr, tb, "hard")), tb
    # deepest fast-path depth with a final block that is exactly full (L = 17 mod 24), every frame length
    # modulo the 4-step history block and the 24-window traceback block
    for nbits in (65, 185, 209, 64, 66, 67, 88, 89, 90, 91):
        _, xx = helpers.channel_frames(tr, rs, 5, nbits, "hard", "cont", flip=0.08)
        for tb in (None, 46, 48, 7, 8):
            assert np.array_equal(viterbi_decode_batch(xx.astype(np.uint8), tr, tb, "hard"),
                                  oracle.viterbi_decode_batch(xx, tr, tb, "hard")), (nbits, tb)
    big = torch.from_numpy(np.concatenate([np.zeros((9, 3)), x], axis=1).astype(np.uint8)).cuda()
    view = big[:, 3:]                                # non-contiguous view: the wrapper must densify it
    assert np.array_equal(viterbi_decode_batch(view, tr, None, "hard").cpu().numpy(),
                          oracle.viterbi_decode_batch(x, tr, None, "hard"))


@pytest.mark.parametrize("mode", ["soft", "unquantized"])
def test_float_decode_is_batch_invariant(mode):
    """a frame's decode must not depend on what it is batched with.  (a) one outlier frame (huge
    values, +-inf) in the batch leaves every other frame's output unchanged; (b) frames [0, B) in one call equal two calls
    of B/2; (c) the outlier frame itself still decodes like the oracle."""
    import torch
    tr = helpers.k7()
    rs = np.random.RandomState(21)
    msgs, x = helpers.channel_frames(tr, rs, 96, 512, mode, "cont", ebn0_db=3.0)
    x = x.astype(np.float32)
    base = viterbi_decode_batch(x, tr, None, mode)
    xo = x.copy()
    xo[17] *= 1.0e6
    if mode == "soft":
        xo[40, ::7] = np.inf
        xo[40, 3::7] = -np.inf
    got = viterbi_decode_batch(xo, tr, None, mode)
    keep = np.ones(96, bool)
    keep[[17, 40]] = False
    assert np.array_equal(got[keep], base[keep])
    want = oracle.viterbi_decode_batch(xo[[17, 40]].astype(np.float64), tr, None, mode)
    assert (got[[17, 40]] == want).mean() >= 0.995
    xt = torch.from_numpy(x).cuda()
    full = viterbi_decode_batch(xt, tr, None, mode)
    h0 = viterbi_decode_batch(xt[:48].contiguous(), tr, None, mode)
    h1 = viterbi_decode_batch(xt[48:].contiguous(), tr, None, mode)
    assert torch.equal(full, torch.cat([h0, h1]))
    assert np.array_equal(full.cpu().numpy(), base)


def test_c2_shape_soft_vs_oracle():
    """BASELINE config 2 shape: K=7 soft-decision, N=4096, AWGN Eb/N0 = 4 dB: 64 frames against the fp64 oracle."""
    tr = helpers.k7()
    rs = np.random.RandomState(22)
    msgs, x = helpers.channel_frames(tr, rs, 64, 4096, "soft", "cont", ebn0_db=4.0)
    want = oracle.viterbi_decode_batch(x, tr, None, "soft", threads=8)
    got = viterbi_decode_batch(x.astype(np.float32), tr, None, "soft")
    assert (got != want).mean() <= 1e-4
    assert abs(int((got != msgs).sum()) - int((want != msgs).sum())) <= 8


def test_packed_hard_equals_unpacked():
    """cpb_viterbi_decode_packed / _host_packed: 1 bit per bit in and out (numpy.packbits order) gives exactly the bits of
    the byte-per-bit path, on the device and through the host pipeline, for every fast-path code and several depths."""
    import torch
    rs = np.random.RandomState(31)
    for make in (helpers.k7, helpers.k7_wifi_quirk):
        tr = make()
        for nbits, term in ((1024, "cont"), (120, "cont"), (250, "term")):
            _, x = helpers.channel_frames(tr, rs, 77, nbits, "hard", term, flip=0.05)
            x = x.astype(np.uint8)
            if x.shape[1] % 16:
                continue
            for tb in (None, 30, 10, 46):
                want = oracle.viterbi_decode_batch(x.astype(np.float64), tr, tb, "hard", threads=4)
                xp = np.packbits(x, axis=1)
                got_h = viterbi_decode_batch(xp, tr, tb, "hard", packed=True)
                assert got_h.shape == (77, want.shape[1] // 8)
                assert np.array_equal(np.unpackbits(got_h, axis=1), want), (nbits, tb)
                got_d = viterbi_decode_batch(torch.from_numpy(xp).cuda(), tr, tb, "hard", packed=True)
                assert np.array_equal(got_d.cpu().numpy(), got_h)
    with pytest.raises(NotImplementedError):
        viterbi_decode_batch(np.packbits(x, axis=1), tr, 31, "hard", packed=True)
    with pytest.raises(ValueError):
        viterbi_decode_batch(np.zeros((4, 2048), np.uint8), helpers.k7(), None, "hard", out=np.zeros((4, 100), np.uint8))
