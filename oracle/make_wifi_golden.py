"""BER of the UNMODIFIED reference's Wifi80211.link_performance (commpy/wifi80211.py:132-216) for the punctured MCS the GPU
link reproduces batched: generated in the build container, replayed by tests/test_links.py against
Wifi80211.link_performance_gpu (same MCS, same SNR definition, far more bits).

    python oracle/make_wifi_golden.py            # writes tests/golden/wifi_ber.npz  (a few minutes: the reference's 256-QAM
                                                 # soft demapper costs ~3 ms per symbol)
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
CASES = {2: [7.5, 8.5], 8: [25.5, 26.5], 9: [25.0, 26.0]}      # mcs -> SNR_dB points inside the (very steep) waterfall
FRAMES, CHUNK = 64, 600


def main():
    refimport.import_reference()
    import importlib
    rw = importlib.import_module("commpy.wifi80211")
    rch = importlib.import_module("commpy.channels")
    out = {}
    for mcs, snrs in CASES.items():
        bers, per_frame = [], []
        for snr in snrs:                       # one call per point: the reference abandons a sweep after a point that
            np.random.seed(1000 + mcs)         # ends below err_min errors (links.py:340-341)
            w = rw.Wifi80211(mcs)
            ch = rch.SISOFlatChannel(None, (1 + 0j, 0j))
            b, bes, ces, ncs = w.link_performance(ch, [snr], FRAMES, 10 ** 9, CHUNK, stop_on_surpass_error=False)
            bers.append(float(b[0]))
            per_frame.append(np.asarray(bes[0], dtype=np.int64).reshape(-1))
            print(mcs, snr, b, per_frame[-1].sum(), len(per_frame[-1]), flush=True)
        out["mcs%d_snr" % mcs] = np.array(snrs)
        out["mcs%d_ber" % mcs] = np.array(bers, dtype=np.float64)
        out["mcs%d_frame_errors" % mcs] = np.stack(per_frame)          # bit errors of every transmission (frame of CHUNK bits)
    out["chunk"] = np.array(CHUNK)
    np.savez_compressed(os.path.join(GOLD, "wifi_ber.npz"), **out)


if __name__ == "__main__":
    main()
