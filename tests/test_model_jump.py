"""CPU check of the ALGORITHM of the K=7 fast Viterbi kernel (history keys + jump traceback): the NumPy model in
tests/model_jump_viterbi.py against the oracle, for both traceback formulations (one walk per output bit = the
definition; the kernel's block traceback with retired paths)."""
import numpy as np
import pytest

import helpers
import model_jump_viterbi as mj
from oracle import oracle


@pytest.mark.parametrize("make,gens", [(helpers.k7, (0o133, 0o171)), (helpers.k7_wifi_quirk, (5, 43))])
def test_model_matches_oracle(make, gens):
    tr = make()
    rs = np.random.RandomState(5)
    for nbits, flip, term in ((120, 0.12, "cont"), (65, 0.1, "cont"), (209, 0.08, "term")):
        _, x = helpers.channel_frames(tr, rs, 2, nbits, "hard", term, flip=flip)
        for D in (None, 7, 9, 15, 46):
            want = oracle.viterbi_decode_batch(x, tr, D, "hard")
            for f in range(x.shape[0]):
                assert np.array_equal(mj.decode(x[f], gens[0], gens[1], D, simple=True), want[f]), (nbits, D)
                assert np.array_equal(mj.decode(x[f], gens[0], gens[1], D), want[f]), (nbits, D)
