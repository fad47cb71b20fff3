"""Outputs of the UNMODIFIED reference's MIMO detectors (commpy/modulation.py:325-646: kbest, best_first_detector,
max_log_approx, bit_lvl_repr) and channels (commpy/channels.py: SISOFlatChannel, MIMOFlatChannel) on seeded inputs, replayed
by tests/test_mimo.py against the host-side mirrors in commpy_b200.modulation / commpy_b200.channels.

    python oracle/make_mimo_golden.py            # writes tests/golden/mimo.npz (seconds)
TEST INFRASTRUCTURE ONLY."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != HERE]
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import numpy as np

import refimport

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
N_CASES = 24


def draw_case(rng, m, nt, nr, snr_db):
    """one transmission: (bits, symbols, H, y, noise_var) for an m-QAM nt x nr uncorrelated Rayleigh channel"""
    nb = int(np.log2(m))
    bits = rng.integers(0, 2, nt * nb)
    h = (rng.standard_normal((nr, nt)) + 1j * rng.standard_normal((nr, nt))) * np.sqrt(0.5)
    noise_var = nt / 10 ** (snr_db / 10)
    noise = (rng.standard_normal(nr) + 1j * rng.standard_normal(nr)) * np.sqrt(noise_var / 2)
    return bits, h, noise, noise_var


def coded_link(rm, rch):
    """the reference's own coded-MIMO acceptance case (commpy/tests/test_links.py:61-86) run with a fixed seed: ~100 s"""
    import importlib
    rl = importlib.import_module("commpy.links")
    rld = importlib.import_module("commpy.channelcoding.ldpc")
    np.random.seed(17)
    qam = rm.QAMModem(16)
    ch = rch.MIMOFlatChannel(4, 4)
    ch.uncorr_rayleigh_fading(complex)
    params = rld.get_ldpc_code_params(os.path.join(refimport.REF_ROOT, "commpy/channelcoding/designs/ldpc/wimax/1440.720.txt"), True)
    model = rl.LinkModel(
        lambda bits: qam.modulate(rld.triang_ldpc_systematic_encode(bits, params, False).reshape(-1, order="F")), ch,
        lambda y, h, c, nv: rm.best_first_detector(y, h, c, (1, 3, 5), nv, lambda s: qam.demodulate(s, "hard"), 500),
        qam.num_bits_symbol, qam.constellation, qam.Es,
        lambda llrs: rld.ldpc_bp_decode(llrs, params, "MSA", 15)[0][:720].reshape(-1, order="F"), 0.5)
    return rl.link_performance(model, np.arange(17, 20, 1), 5e5, 200, 720, model.rate)


def main():
    refimport.import_reference()
    import importlib
    rm = importlib.import_module("commpy.modulation")
    rch = importlib.import_module("commpy.channels")
    out = {}
    rng = np.random.default_rng(2024)
    for m, nt, nr in ((4, 2, 2), (16, 4, 4), (16, 2, 4), (64, 3, 3)):
        modem = rm.QAMModem(m)
        tag = "q%d_%dx%d" % (m, nt, nr)
        H, Y, NV, KH, KS, BF, ML = [], [], [], [], [], [], []
        for c in range(N_CASES):
            bits, h, noise, nv = draw_case(rng, m, nt, nr, rng.uniform(5, 25))
            x = modem.modulate(bits)
            y = h.dot(x) + noise
            H.append(h), Y.append(y), NV.append(nv)
            KH.append(rm.kbest(y, h, modem.constellation, 8))
            KS.append(rm.kbest(y, h, modem.constellation, 8, nv, "soft", lambda s: modem.demodulate(s, "hard")))
            if nt == nr:                     # the reference's best-first detector indexes its stacks by h.shape[0]
                BF.append(rm.best_first_detector(y, h, modem.constellation, (1,) + (2,) * (nt - 1) if nt > 1 else (1,), nv,
                                                 lambda s: modem.demodulate(s, "hard"), 500))
            pts = modem.constellation[rng.integers(0, m, (nt, 12))]
            ML.append(np.concatenate([pts.reshape(-1), rm.max_log_approx(y, h, nv, pts, lambda s: modem.demodulate(s, "hard"))]))
        out[tag + "_h"], out[tag + "_y"], out[tag + "_nv"] = np.array(H), np.array(Y), np.array(NV)
        out[tag + "_kbest_hard"], out[tag + "_kbest_soft"] = np.array(KH), np.array(KS)
        if BF:
            out[tag + "_best_first"] = np.array(BF)
        out[tag + "_maxlog"] = np.array(ML)
    w = np.array([2.0, 1.0, 0.5, 0.25])
    hb = rng.standard_normal((3, 2))
    out["blr_h"], out["blr_w"], out["blr_out"] = hb, w, rm.bit_lvl_repr(hb, w)
    # channels: seeded propagation (global numpy RNG, as the reference uses)
    msg = (rng.standard_normal(37) + 1j * rng.standard_normal(37))
    for name, ch in (("siso_awgn", rch.SISOFlatChannel(None, (1 + 0j, 0j))),
                     ("siso_rice", rch.SISOFlatChannel(None, (0.6 + 0j, 0.64))),
                     ("mimo_2x3", rch.MIMOFlatChannel(2, 3)),
                     ("mimo_4x4", rch.MIMOFlatChannel(4, 4))):
        if name.startswith("mimo"):
            ch.uncorr_rayleigh_fading(complex)
        ch.set_SNR_dB(7.0, 0.5, 1.3)
        np.random.seed(99)
        out["ch_" + name + "_out"] = ch.propagate(msg)
        out["ch_" + name + "_gains"] = ch.channel_gains
        out["ch_" + name + "_noise_std"] = np.array(ch.noise_std)
    out["ch_msg"] = msg
    # correlated MIMO (Kronecker model with exponential correlation matrices)
    ch = rch.MIMOFlatChannel(3, 3)
    t = 0.5 * np.exp(0.3j)
    rt = np.array([[t ** (j - i) if j >= i else np.conj(t ** (i - j)) for j in range(3)] for i in range(3)])
    ch.fading_param = (np.zeros((3, 3), complex), rt, rt.conj())
    ch.set_SNR_dB(10.0)
    np.random.seed(7)
    out["ch_mimo_corr_out"] = ch.propagate(msg)
    out["ch_mimo_corr_gains"] = ch.channel_gains
    out["ch_mimo_corr_rt"] = rt
    out["link_best_first_ldpc_ber"] = coded_link(rm, rch)
    np.savez_compressed(os.path.join(GOLD, "mimo.npz"), **out)
    print("wrote", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
