"""Link-level tests: host loop (CPU) and the batched GPU chain (config C5 shape)."""
import math

import numpy as np
import pytest

import helpers
from commpy_b200.links import AwgnSisoChannel, ConvLinkGPU, LinkModel, _ff_taps, link_performance


def _q(x):
    return 0.5 * math.erfc(x / math.sqrt(2))


def test_linkmodel_host_loop_bpsk_matches_theory():
    """commpy/tests/test_links.py:37-42 in spirit: uncoded antipodal signalling over real AWGN vs the Q function."""
    np.random.seed(17121996)
    ch = AwgnSisoChannel(is_complex=False, rng=np.random.RandomState(1))
    model = LinkModel(lambda b: 2.0 * np.asarray(b) - 1.0, ch, lambda y, h, c, nv: (y > 0).astype(int), 1,
                      np.array([-1.0, 1.0]), Es=1.0)
    snrs = np.array([0.0, 4.0, 6.0])
    bers = link_performance(model, snrs, 4e5, 600, 2000)
    want = [_q(math.sqrt(10 ** (s / 10))) for s in snrs]       # noise_std^2 = Es/10^(SNR/10)
    assert np.allclose(bers, want, rtol=0.25)
    # arity sniffing (links.py:306): a 6-argument decoder is called with the full argument list
    seen = {}

    def dec6(y, h, c, nv, arr, bps):
        seen["n"] = bps
        return arr
    model6 = LinkModel(lambda b: 2.0 * np.asarray(b) - 1.0, ch, lambda y, h, c, nv: (y > 0).astype(int), 1,
                       np.array([-1.0, 1.0]), Es=1.0, decoder=dec6)
    model6.link_performance(np.array([3.0]), 2000, 10, 1000)
    assert seen["n"] == 1


def test_linkmodel_reproduces_reference_bers_for_the_same_seed():
    """`LinkModel.link_performance` + `AwgnSisoChannel` consume numpy's global random stream exactly like the reference's
    loop (links.py:313-337) with `SISOFlatChannel(None, (1 + 0j, 0j))` (channels.py:181-221), so a seeded run gives the
    reference's BERs to the last digit.  Expected values: the reference itself, run in the build container with
    np.random.seed(seed), PSKModem(4), SNRs (0, 4, 8) dB, send_max 20000, err_min 200, send_chunk 600 and a nearest-point
    hard receiver."""
    from commpy_b200.modulation import PSKModem
    expected = {5: (303 / 1800, 223 / 3600, 135 / 20400), 6: (274 / 1800, 207 / 3600, 146 / 20400)}
    for seed, want in expected.items():
        np.random.seed(seed)
        modem = PSKModem(4)
        cst = np.asarray(modem.constellation)
        nb = modem.num_bits_symbol

        def receive(y, H, constellation, noise_var):
            idx = np.abs(np.asarray(y)[:, None] - cst[None, :]).argmin(1)
            return ((idx[:, None] >> np.arange(nb - 1, -1, -1)) & 1).reshape(-1)

        model = LinkModel(modem.modulate, AwgnSisoChannel(True), receive, nb, modem.constellation, modem.Es)
        bers = link_performance(model, np.arange(0, 9, 4), 20000, 200, 600)
        assert np.allclose(bers, want, rtol=0, atol=1e-12), (seed, bers)


def test_full_metrics_loop_reproduces_reference_for_the_same_seed():
    """`LinkModel.link_performance_full_metrics` (links.py:155-267) with np.random.seed(11), PSKModem(4), SNRs (0, 4, 8) dB,
    tx_max 40, err_min 100, send_chunk 600: bit errors per transmission exactly as the reference produced them in the
    build container."""
    from commpy_b200.modulation import PSKModem
    np.random.seed(11)
    modem = PSKModem(4)
    cst = np.asarray(modem.constellation)

    def receive(y, H, constellation, noise_var):
        idx = np.abs(np.asarray(y)[:, None] - cst[None, :]).argmin(1)
        return ((idx[:, None] >> np.arange(1, -1, -1)) & 1).reshape(-1)

    model = LinkModel(modem.modulate, AwgnSisoChannel(True), receive, 2, modem.constellation, modem.Es)
    BERs, BEs, CEs, NCs = model.link_performance_full_metrics(np.arange(0, 9, 4), 40, 100, 600)
    want = [[95, 88], [41, 23, 30, 33], [2, 5, 1, 5, 6, 5, 5, 4, 3, 6, 2, 4, 2, 3, 2, 3, 6, 9, 1, 3, 7, 4, 1, 3, 4, 5]]
    for i, row in enumerate(want):
        assert list(np.asarray(BEs)[i, :len(row)]) == row and not np.asarray(BEs)[i, len(row):].any()
        assert np.asarray(NCs)[i].sum() == len(row)
        assert abs(BERs[i] - sum(row) / (600 * len(row))) < 1e-12


def test_ff_taps_and_chunk_rounding():
    t = _ff_taps(helpers.k7())
    assert t.shape == (2, 7)
    assert int("".join(str(b) for b in t[0][::-1]), 2) == 0o133 and int("".join(str(b) for b in t[1][::-1]), 2) == 0o171
    assert _ff_taps(helpers.rsc_k4()) is None          # recursive code: no tap form
    # send_chunk is rounded down to a multiple of denominator(1/(bits*nb_tx)/rate) (links.py:302-303):
    # 6 bits/symbol at rate 3/4 -> (1/6)/(3/4) = 2/9 -> chunks of 36 instead of 37
    calls = []

    def modulate(bits):
        calls.append(len(bits))
        return np.zeros(8, complex)
    m = LinkModel(modulate, AwgnSisoChannel(rng=np.random.RandomState(0)), lambda y, h, c, nv: np.zeros(64, int), 6, None)
    m.link_performance(np.array([0.0]), 100, 10 ** 9, 37, 0.75)
    assert calls and set(calls) == {36}


@pytest.mark.gpu
def test_conv_link_gpu_qpsk_k7_soft_ber_and_oracle_agreement():
    import torch
    from oracle import oracle
    from commpy_b200.modulation import QAMModem
    tr = helpers.k7()
    link = ConvLinkGPU(tr, QAMModem(4), frame_bits=1024, frames_per_batch=512, decoding_type="soft", seed=5)
    ebn0 = np.array([2.0, 3.0])
    snr = ebn0 + 10 * math.log10(2)          # SNR = Eb/N0 + 10 log10(bits/symbol); set_SNR_dB divides by the rate itself
    bers = link.link_performance(snr, send_max=2e6, err_min=400)
    # K=7 soft Viterbi with a 30-step window: ~9e-3 at 2 dB and ~7e-4 at 3 dB; loose statistical bounds
    assert 2e-3 < bers[0] < 3e-2 and 1e-4 < bers[1] < 3e-3, bers
    # the decoded bits of one batch agree with the CPU oracle fed the same LLRs
    msg, y, nv = link.make_batch(float(snr[0]), 0, torch)
    llr = link.modem.demodulate_batch(y, "soft", nv)
    cnt = torch.zeros(3, dtype=torch.int64, device="cuda")
    dec = link.receive_decode_count(msg, y, nv, cnt, torch).cpu().numpy()
    want = oracle.viterbi_decode_batch(llr[:24].cpu().numpy().astype(np.float64), tr, None, "soft", threads=4)
    assert (dec[:24] == want).mean() > 0.9999
    assert int(cnt[0]) == int((dec != msg.cpu().numpy()).sum())


@pytest.mark.gpu
def test_link_performance_point_overlap_changes_nothing():
    """ConvLinkGPU.link_performance reads a point's last counters after the next point's first batch has been issued
    (overlap_points=True, the default): BERs and counters equal the read-at-once order, for single- and multi-batch points,
    with the reference's early end of the sweep (links.py:339-341) and without it."""
    from commpy_b200.modulation import QAMModem
    tr = helpers.k7()
    snr = np.array([1.0, 2.5, 4.0, 7.0, 9.0]) + 10 * math.log10(2)
    for frames, send_max, err_min, stop_early in ((256, 2e5, 300, True), (256, 1.2e6, 2000, True), (128, 3e5, 10 ** 9, False),
                                                  (256, 6e5, 150, False)):
        res = []
        for overlap in (False, True):
            link = ConvLinkGPU(tr, QAMModem(4), frame_bits=1024, frames_per_batch=frames, decoding_type="soft", seed=9)
            bers, cnt = link.link_performance(snr, send_max=send_max, err_min=err_min, return_counters=True,
                                              stop_early=stop_early, overlap_points=overlap)
            res.append((bers, cnt.cpu().numpy()))
        assert np.array_equal(res[0][0], res[1][0]), (frames, send_max, err_min, res)
        assert np.array_equal(res[0][1], res[1][1])
        assert res[0][0][0] > 0
        if stop_early:
            assert res[0][0][-1] == 0.0            # the sweep ended before the last point, which stays unfilled


def test_philox_model_known_answers():
    """The NumPy Philox4x32-10 the TX-chain test is built on reproduces the Random123 known-answer vectors."""
    r = helpers.philox4x32_10(np.array([[0, 0, 0, 0]], dtype=np.uint64), (0, 0))[0]
    assert [int(x) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = helpers.philox4x32_10(np.array([[0xffffffff] * 4], dtype=np.uint64), (0xffffffff, 0xffffffff))[0]
    assert [int(x) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = helpers.philox4x32_10(np.array([[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], dtype=np.uint64),
                              (0xa4093822, 0x299f31d0))[0]
    assert [int(x) for x in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.gpu
@pytest.mark.parametrize("modem_m,frame_bits", [(4, 200), (16, 1024), (64, 300), (256, 4096)])
def test_conv_link_tx_kernel_matches_numpy_model(modem_m, frame_bits):
    """cpb_conv_link_tx: message bits and noiseless symbols bit-exact against conv_encode + Modem.modulate on the
    Philox message stream, noise equal to the float64 Box-Muller model to fast-math accuracy, and the frames do not
    depend on how they are split over calls (first_frame) -- the multi-GPU sharding contract."""
    import torch
    from commpy_b200.links import conv_link_tx
    from commpy_b200.modulation import QAMModem
    tr = helpers.k7()
    modem = QAMModem(modem_m)
    frames, seed, first = 5, 0x1234567890abcdef, (1 << 33) + 7
    want_msg, want_y0 = helpers.conv_link_tx_model(tr, modem, frames, frame_bits, seed, first, 0.0)
    msg, y0 = conv_link_tx(tr, modem, frames, frame_bits, seed, first, 0.0)
    assert np.array_equal(msg.cpu().numpy(), want_msg)
    assert np.array_equal(y0.cpu().numpy(), want_y0.astype(np.complex64))
    _, want_y = helpers.conv_link_tx_model(tr, modem, frames, frame_bits, seed, first, 0.75)
    msg2, y = conv_link_tx(tr, modem, frames, frame_bits, seed, first, 0.75)
    assert torch.equal(msg, msg2)
    assert np.abs(y.cpu().numpy() - want_y).max() < 2e-3
    # split invariance: frames [first+2, first+5) generated on their own are the same frames
    msg3, y3 = conv_link_tx(tr, modem, 3, frame_bits, seed, first + 2, 0.75)
    assert torch.equal(msg3, msg[2:]) and torch.equal(y3, y[2:])
    # noise statistics over a larger batch: unit variance per component, zero mean, uncorrelated I/Q
    _, yb = conv_link_tx(tr, modem, 256, frame_bits, 7, 0, 1.0)
    _, yc = conv_link_tx(tr, modem, 256, frame_bits, 7, 0, 0.0)
    z = (yb - yc).cpu().numpy().reshape(-1)
    nz = z.size
    assert abs(z.real.mean()) < 5 / math.sqrt(nz) and abs(z.imag.mean()) < 5 / math.sqrt(nz)
    assert abs(z.real.var() - 1) < 8 / math.sqrt(nz) and abs(z.imag.var() - 1) < 8 / math.sqrt(nz)
    assert abs((z.real * z.imag).mean()) < 5 / math.sqrt(nz)
    # recursive trellises are refused (no tap form)
    with pytest.raises(NotImplementedError):
        conv_link_tx(helpers.rsc_k4(), QAMModem(4), 2, 64, 1, 0, 0.1)


@pytest.mark.gpu
@pytest.mark.parametrize("modem_m", [4, 16, 256])
def test_conv_link_tx_word_parallel_kernel_equals_bit_serial(modem_m):
    """The word-parallel TX kernel (n = 2, no puncturing, 2 / 4 / 8 bits per symbol, 128-bit message blocks) and the bit-serial
    one (CPB_OPT_TX_FORCE_GENERIC) produce the same message bytes and the same received symbols, noise included."""
    import torch
    from commpy_b200 import _lib
    from commpy_b200.links import conv_link_tx
    from commpy_b200.modulation import QAMModem
    modem = QAMModem(modem_m)
    from commpy_b200.channelcoding import Trellis
    k3 = Trellis(np.array([2]), np.array([[5, 7]]))
    for tr, frame_bits in ((helpers.k7(), 1024), (helpers.k7_wifi_quirk(), 128), (helpers.k7(), 4096), (k3, 256)):
        out = {}
        for force in (1, 0):
            _lib.set_option(_lib.OPT_TX_FORCE_GENERIC, force)
            try:
                out[force] = conv_link_tx(tr, modem, 37, frame_bits, 0xfeedbeef12345, (1 << 32) - 5, 0.6)
            finally:
                _lib.set_option(_lib.OPT_TX_FORCE_GENERIC, 0)
        assert torch.equal(out[0][0], out[1][0])
        assert torch.equal(torch.view_as_real(out[0][1]), torch.view_as_real(out[1][1]))


@pytest.mark.gpu
def test_dropin_linkmodel_with_gpu_receiver_and_decoder():
    """commpy/examples/conv_encode_decode.py:99-127 shape: QPSK, (5,7) code, hard and unquantized Viterbi via LinkModel."""
    from commpy_b200.channelcoding import Trellis, conv_encode, viterbi_decode
    from commpy_b200.modulation import QAMModem
    np.random.seed(3)
    tr = Trellis(np.array([2]), np.array([[5, 7]]))
    modem = QAMModem(4)
    ch = AwgnSisoChannel(rng=np.random.RandomState(4))

    def modulate(bits):
        return modem.modulate(conv_encode(bits, tr, "cont"))

    def receiver_hard(y, h, constellation, noise_var):
        return modem.demodulate(y, "hard")

    def decoder_hard(msg):
        return viterbi_decode(msg, tr)

    model = LinkModel(modulate, ch, receiver_hard, modem.num_bits_symbol, modem.constellation, modem.Es, decoder_hard, 0.5)
    ber = model.link_performance(np.array([5.0]) + 10 * math.log10(2), 40000, 200, 1000, 0.5)[0]
    assert 1e-4 < ber < 4e-3, ber                # commpy/channelcoding/README.md:159-160 quotes 7.8e-4 (hard, 5 dB)


def test_wifi80211_tables_match_reference():
    from commpy_b200.wifi80211 import Wifi80211
    assert [Wifi80211(m)._get_coding() for m in range(10)] == [(1, 2), (1, 2), (3, 4), (1, 2), (3, 4), (2, 3), (3, 4),
                                                              (5, 6), (3, 4), (5, 6)]
    assert [Wifi80211(m).get_modem().m for m in range(10)] == [2, 4, 4, 16, 16, 64, 64, 64, 256, 256]
    assert Wifi80211._get_puncture_matrix(3, 4) == [1, 1, 1, 0, 0, 1] and Wifi80211._get_puncture_matrix(1, 2) is None
    tr = Wifi80211._get_trellis()            # decimal (133,171) quirk: taps (5, 43)
    assert np.array_equal(tr.output_table, helpers.k7_wifi_quirk().output_table)


@pytest.mark.gpu
def test_wifi80211_link_performance_runs_on_gpu_path():
    """Wifi80211(mcs).link_performance end to end (punctured 3/4 and unpunctured), BER falls with SNR."""
    from commpy_b200.wifi80211 import Wifi80211
    np.random.seed(11)
    for mcs, snrs in ((1, np.array([4.0, 12.0])), (4, np.array([10.0, 22.0]))):
        w = Wifi80211(mcs)
        ch = AwgnSisoChannel(rng=np.random.RandomState(12))
        BERs, BEs, CEs, NCs = w.link_performance(ch, snrs, 12, 1, 600, stop_on_surpass_error=False)
        assert BERs.shape == (2,) and BEs.shape == (2, 12) and NCs.sum() == 24
        assert BERs[0] > BERs[1] and BERs[1] < 0.05, (mcs, BERs)


@pytest.mark.gpu
def test_published_ber_of_reference_readme():
    """The experiment of commpy/channelcoding/README.md:81-163 (QPSK, Trellis([6], [[5, 7]]), tb_depth = 5*(m+1) = 35,
    Eb/N0 = 5 dB, 10 x 1e5 bits) on the GPU path.  The README prints 2.8e-5 / 7.81e-4 / 9.064e-3 (unquantized / hard /
    uncoded); its uncoded figure is 0.5 dB off the QPSK closed form, i.e. those numbers come from an older noise
    convention and are not what the CURRENT reference code produces.  What is asserted here: the uncoded BER equals
    theory, the hard stream decodes bit-exactly like the oracle (= the current reference), and soft beats hard."""
    from commpy_b200.channelcoding import Trellis, viterbi_decode_batch
    from commpy_b200.modulation import PSKModem
    rs = np.random.RandomState(2024)
    N, trials, M = 100000, 10, 4
    k, rate = 2, 0.5
    trellis = Trellis(np.array([6]), np.array([[5, 7]]))
    modem = PSKModem(M)
    tb_depth = 5 * (6 + 1)
    EbNo = 5
    snrdB = EbNo + 10 * np.log10(k * rate)
    noiseVar = 10 ** (-snrdB / 10)
    msgs = rs.randint(0, 2, (trials, N))
    coded = helpers.encode_batch(msgs, trellis, "term")
    soft_in, hard_in, unc_err = [], [], 0
    for t in range(trials):
        x = modem.modulate(coded[t])
        xu = modem.modulate(msgs[t])
        Es = np.mean(np.abs(x) ** 2)
        No = Es / ((10 ** (EbNo / 10)) * np.log2(M))
        noisy = x + np.sqrt(No / 2) * (rs.randn(len(x)) + 1j * rs.randn(len(x)))
        noisy_u = xu + np.sqrt(No / 2) * (rs.randn(len(xu)) + 1j * rs.randn(len(xu)))
        soft_in.append(modem.demodulate(noisy, "soft", noiseVar))
        hard_in.append(modem.demodulate(noisy, "hard"))
        unc_err += int((modem.demodulate(noisy_u, "hard") != msgs[t]).sum())
    dec_soft = viterbi_decode_batch(np.stack(soft_in).astype(np.float32), trellis, tb_depth, "unquantized")
    dec_hard = viterbi_decode_batch(np.stack(hard_in).astype(np.uint8), trellis, tb_depth, "hard")
    ber_soft = (dec_soft[:, :N] != msgs).mean()
    ber_hard = (dec_hard[:, :N] != msgs).mean()
    ber_unc = unc_err / (trials * N)
    print("README experiment on the GPU path: uncoded %.3e hard %.3e unquantized %.3e" % (ber_unc, ber_hard, ber_soft))
    assert abs(ber_unc - _q(math.sqrt(2 * 10 ** 0.5))) / _q(math.sqrt(2 * 10 ** 0.5)) < 0.05, ber_unc     # QPSK theory
    assert ber_hard < 2e-4 and ber_soft <= ber_hard, (ber_hard, ber_soft)
    # and bit-exact with the oracle on the hard stream of one trial
    from oracle import oracle
    want = oracle.viterbi_decode(hard_in[0].astype(np.float64), trellis, tb_depth, "hard")
    assert np.array_equal(dec_hard[0], want)


@pytest.mark.gpu
def test_counters_identical_for_1_2_4_8_ranks(monkeypatch):
    """SURVEY 4(4) / VERDICT r1: same seed => same error counters whatever the number of ranks.  The rank layout is emulated
    in one process (RANK / WORLD_SIZE in the environment; no process group, so the all-reduce is the sum taken here): the
    same 1,024 global frames per step are generated and decoded as 1, 2, 4 and 8 shards."""
    import torch
    from commpy_b200.modulation import QAMModem
    tr = helpers.k7()
    snr = 11.0 + 10 * np.log10(8)
    totals = {}
    for world in (1, 2, 4, 8):
        link = ConvLinkGPU(tr, QAMModem(256), frame_bits=1024, frames_per_batch=1024 // world, decoding_type="soft", seed=77)
        tot = torch.zeros(3, dtype=torch.int64, device="cuda")
        for rank in range(world):
            monkeypatch.setenv("RANK", str(rank))
            monkeypatch.setenv("WORLD_SIZE", str(world))
            for b in range(3):
                msg, y, nv = link.make_batch(snr, b, torch)
                link.receive_decode_count(msg, y, nv, tot, torch)
        totals[world] = tot.cpu().numpy().copy()
    monkeypatch.delenv("RANK")
    monkeypatch.delenv("WORLD_SIZE")
    assert totals[1][0] > 0, "the test point must have bit errors to compare"
    for world in (2, 4, 8):
        assert np.array_equal(totals[world], totals[1]), (world, totals)


@pytest.mark.gpu
def test_c5_chain_256qam_soft_vs_oracle():
    """BASELINE config 5 chain at its own shape: 256-QAM symbols of 4096-bit K=7 frames -> soft demapper -> soft Viterbi,
    against the oracle's demapper feeding the oracle's Viterbi on the same received symbols."""
    import torch
    from oracle import oracle
    from commpy_b200.modulation import QAMModem
    tr = helpers.k7()
    link = ConvLinkGPU(tr, QAMModem(256), frame_bits=4096, frames_per_batch=64, decoding_type="soft", seed=5)
    bad = tot = 0
    for ebn0 in (10.0, 13.0):
        msg, y, nv = link.make_batch(ebn0 + 10 * np.log10(8), 0, torch)
        cnt = torch.zeros(3, dtype=torch.int64, device="cuda")
        dec = link.receive_decode_count(msg, y, nv, cnt, torch)[:12].cpu().numpy()
        yy = y[:12].cpu().numpy().astype(np.complex128)
        for f in range(12):
            llr = oracle.demodulate(link.modem, yy[f], "soft", nv)
            want = oracle.viterbi_decode(llr, tr, None, "soft")
            bad += int((dec[f] != want).sum())
            tot += want.size
    assert bad / tot <= 1e-4, (bad, tot)


WIFI_PATTERNS = {"2/3": [1, 1, 1, 0], "3/4": [1, 1, 1, 0, 0, 1], "5/6": [1, 1, 1, 0, 0, 1, 1, 0, 0, 1]}


@pytest.mark.gpu
@pytest.mark.parametrize("rate", sorted(WIFI_PATTERNS))
def test_punctured_viterbi_kernel_vs_oracle(rate):
    """cpb_viterbi_decode_punctured (depuncturing fused into the kernel's load) against depuncturing() + the oracle's
    viterbi_decode on the same punctured LLRs: both 802.11 trellises (octal K=7 and the reference's decimal-quirk one)."""
    from oracle import oracle
    from commpy_b200.channelcoding import depuncturing, puncturing, viterbi_decode_punctured_batch
    pv = WIFI_PATTERNS[rate]
    rs = np.random.RandomState(41)
    for tr in (helpers.k7(), helpers.k7_wifi_quirk()):
        for nbits in (600, 1020):
            msgs = rs.randint(0, 2, (40, nbits))
            coded = helpers.encode_batch(msgs, tr, "cont")
            y = (2.0 * coded - 1) + 0.45 * rs.randn(*coded.shape)
            llr = 2 * y / 0.45 ** 2
            rows = np.stack([puncturing(r, pv) for r in llr])
            got = viterbi_decode_punctured_batch(rows.astype(np.float32), tr, pv, coded.shape[1]).cpu().numpy()
            dep = np.stack([depuncturing(r, pv, coded.shape[1]) for r in rows])
            want = oracle.viterbi_decode_batch(dep, tr, None, "soft", threads=4)
            assert (got != want).mean() <= 2e-4, (rate, nbits, (got != want).sum())
            assert abs(int((got != msgs).sum()) - int((want != msgs).sum())) <= 10
    with pytest.raises(IndexError):
        viterbi_decode_punctured_batch(rows[:, :-5].astype(np.float32), tr, pv, coded.shape[1])


@pytest.mark.gpu
@pytest.mark.parametrize("modem_m,rate,frame_bits", [(4, "3/4", 1536), (256, "3/4", 1536), (256, "5/6", 1600), (64, "2/3", 1536)])
def test_conv_link_tx_punctured_matches_numpy_model(modem_m, rate, frame_bits):
    """The TX kernel with puncturing between encoder and mapper: message bits and noiseless symbols equal the NumPy model
    (Philox message, conv_encode, puncturing, modulate), whatever the split into calls."""
    import torch
    from commpy_b200.links import conv_link_tx
    from commpy_b200.modulation import QAMModem
    tr = helpers.k7_wifi_quirk()
    modem = QAMModem(modem_m)
    pv = WIFI_PATTERNS[rate]
    msg, y = conv_link_tx(tr, modem, 6, frame_bits, seed=123456789, first_frame=10, noise_sigma=0.0, puncture=pv)
    m_ref, y_ref = helpers.conv_link_tx_model(tr, modem, 6, frame_bits, 123456789, 10, 0.0, puncture=pv)
    assert np.array_equal(msg.cpu().numpy(), m_ref)
    assert np.allclose(y.cpu().numpy(), y_ref, atol=1e-5)
    msg2, y2 = conv_link_tx(tr, modem, 2, frame_bits, seed=123456789, first_frame=13, noise_sigma=0.7, puncture=pv)
    msg3, y3 = conv_link_tx(tr, modem, 6, frame_bits, seed=123456789, first_frame=10, noise_sigma=0.7, puncture=pv)
    assert torch.equal(msg2, msg3[3:5]) and torch.equal(y2, y3[3:5])


@pytest.mark.gpu
@pytest.mark.parametrize("mcs", [2, 8, 9])
def test_wifi80211_gpu_link_ber_matches_reference_golden(mcs):
    """the batched punctured GPU link (Wifi80211.link_performance_gpu) against BERs measured with the
    UNMODIFIED reference's Wifi80211.link_performance (oracle/make_wifi_golden.py -> tests/golden/wifi_ber.npz): same MCS,
    same trellis quirk, same SNR definition; the reference points have a few hundred bit errors (bursty), so the
    comparison allows 40 % plus three binomial sigmas."""
    import os
    from commpy_b200.wifi80211 import Wifi80211
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wifi_ber.npz"))
    snrs, fe, chunk = g["mcs%d_snr" % mcs], g["mcs%d_frame_errors" % mcs].astype(np.float64), int(g["chunk"])
    ref = fe.mean(axis=1) / chunk                                   # BER of the reference at each point
    se = fe.std(axis=1, ddof=1) / chunk / np.sqrt(fe.shape[1])      # its standard error (frames fail as a whole: bursty)
    w = Wifi80211(mcs)
    got = []
    for snr in snrs:                                                # frames of the reference's length: the quirk code propagates errors
        got.append(w.link_performance_gpu([float(snr)], send_max=6e6, err_min=10 ** 9, send_chunk=chunk, frames_per_batch=2048,
                                          seed=3, stop_early=False)[0])
    assert w.gpu_link.frame_bits == chunk
    for b_ref, s_ref, b_gpu in zip(ref, se, got):
        assert abs(b_gpu - b_ref) <= 4 * s_ref + 0.15 * b_ref + 2e-3, (mcs, list(snrs), list(ref), list(se), got)
    assert max(ref) > 0.01, "the golden points must sit inside the waterfall"


@pytest.mark.gpu
def test_turbo_link_tx_kernel_matches_numpy_model_and_link_runs():
    """cpb_turbo_link_tx (device-side turbo_encode + BPSK + AWGN): message bits and the noiseless streams
    equal the host turbo_encode mirror on the Philox message (incl. its unterminated-'rsc' accident), the noise matches the
    float64 Box-Muller model, frames do not depend on the split into calls; the batched turbo link's BER falls with Eb/N0."""
    import torch
    from commpy_b200.channelcoding import RandInterlv
    from commpy_b200.links import TurboLinkGPU, turbo_link_tx
    rsc = helpers.rsc_k4()
    for N in (256, 1000):
        il = RandInterlv(N, 1)
        want = helpers.turbo_link_tx_model(rsc, il, 4, N, 99, (1 << 32) + 3, 0.0)
        got = turbo_link_tx(rsc, il, 4, N, 99, (1 << 32) + 3, 0.0)
        assert np.array_equal(got[0].cpu().numpy(), want[0])
        for a, b in zip(got[1:], want[1:]):
            assert np.array_equal(a.cpu().numpy(), b.astype(np.float32))
        wn = helpers.turbo_link_tx_model(rsc, il, 4, N, 99, (1 << 32) + 3, 0.8)
        gn = turbo_link_tx(rsc, il, 4, N, 99, (1 << 32) + 3, 0.8)
        for a, b in zip(gn[1:], wn[1:]):
            assert np.abs(a.cpu().numpy() - b).max() < 2e-3
        part = turbo_link_tx(rsc, il, 2, N, 99, (1 << 32) + 5, 0.8)
        for a, b in zip(part, gn):
            assert torch.equal(a, b[2:])
    link = TurboLinkGPU(rsc, RandInterlv(1024, 1), 1024, frames_per_batch=256, iterations=4, seed=8)
    bers = link.link_performance([0.0, 1.5], send_max=1e6, err_min=10 ** 9)
    assert bers[0] > 5 * bers[1] and bers[0] > 1e-3, bers
