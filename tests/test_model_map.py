"""CPU check of the block-rescaled MAP arithmetic (tests/model_map_block.py = tpf::map_lin2_kernel of bcjr.cu) against the
fp64 oracle: same tolerance as the GPU parity tests, on ordinary frames and on frames that force the per-step fallback."""
import numpy as np

import helpers
from model_map_block import map_decode_model
from oracle import oracle

MAP_ATOL, MAP_RTOL = 1e-4, 1e-4


def test_block_rescaled_model_matches_oracle_and_takes_the_fallback():
    tr = helpers.rsc_k4()
    rs = np.random.RandomState(78)
    N, batch = 256, 6
    fallbacks = []
    for s2, amp, lamp in helpers.CONTRADICTED_MAP_CASES:
        ys = amp * rs.choice([-1.0, 1.0], (batch, N)) + np.sqrt(s2) * rs.randn(batch, N)
        yp = amp * rs.choice([-1.0, 1.0], (batch, N)) + np.sqrt(s2) * rs.randn(batch, N)
        La = lamp * rs.randn(batch, N)
        for per_step in (True, False):
            L = map_decode_model(ys, yp, tr, s2, La, per_step).astype(np.float64)
            if not per_step:
                fallbacks.append(sum(map_decode_model.fallbacks))
            for b in range(batch):
                Lo, _ = oracle.map_decode(ys[b], yp[b], tr, s2, La[b], "decode")
                ok = np.isfinite(Lo) & (np.abs(Lo) < 40.0)
                d = np.abs(L[b][ok] - Lo[ok])
                assert (d <= MAP_ATOL + MAP_RTOL * np.abs(Lo[ok])).all(), (s2, per_step, float(d.max()))
    assert fallbacks[0] > 50 and fallbacks[1] > 20 and fallbacks[2] == 0, fallbacks


def test_block_rescaled_model_on_code_words():
    from commpy_b200.channelcoding import conv_encode
    tr = helpers.rsc_k4()
    rs = np.random.RandomState(79)
    N, batch = 512, 4
    msgs = rs.randint(0, 2, (batch, N))
    coded = np.stack([conv_encode(m, tr, "cont") for m in msgs])
    for snr_db in (0.0, 4.0):
        s2 = 1.0 / (2 * 0.5 * 10 ** (snr_db / 10))
        ys = 2.0 * coded[:, 0::2] - 1 + np.sqrt(s2) * rs.randn(batch, N)
        yp = 2.0 * coded[:, 1::2] - 1 + np.sqrt(s2) * rs.randn(batch, N)
        La = rs.randn(batch, N)
        L = map_decode_model(ys, yp, tr, s2, La).astype(np.float64)
        for b in range(batch):
            Lo, _ = oracle.map_decode(ys[b], yp[b], tr, s2, La[b], "decode")
            assert (np.abs(L[b] - Lo) <= MAP_ATOL + MAP_RTOL * np.abs(Lo)).all()
