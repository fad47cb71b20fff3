"""CPU: host-side mirror of the reference's descriptors (Trellis, conv_encode, puncturing, interleaver, modems,
LDPC design files) against known-answer tables -- the reference's own (test_convcode.py:23-127,
test_utilities.py:12-13, test_modulation.py:159-174) and goldens produced by the imported reference."""
import os
import warnings
from itertools import product

import numpy as np
import pytest

import helpers
from commpy_b200.channelcoding import (RandInterlv, Trellis, conv_encode, depuncturing, get_ldpc_code_params,
                                        puncturing, turbo_encode)
from commpy_b200.modulation import Modem, PSKModem, QAMModem
from commpy_b200.utilities import bitarray2dec, dec2bitarray, euclid_dist, hamming_dist, signal_power

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _specs():
    a = np.array
    return {
        "t57": (a([2]), a([[5, 7]]), None, "default", "MSB"),
        "rsc_legacy": (a([2]), a([[1, 7]]), 5, "rsc", "MSB"),
        "r23": (a([2, 1]), a([[5, 7, 0], [0, 2, 3]]), None, "default", "MSB"),
        "r23_lsb": (a([2, 1]), a([[5, 7, 0], [0, 2, 6]]), None, "default", "LSB"),
        "r23_rsc": (a([1, 1]), a([[1, 0, 0], [0, 1, 3]]), a([[2, 2], [3, 1]]), "rsc", "MSB"),
        "k7": (a([6]), a([[0o133, 0o171]]), None, "default", "MSB"),
        "k7_wifi_quirk": (a([6]), a([[133, 171]]), None, "default", "MSB"),
        "rsc_k4": (a([3]), a([[1, 0o15]]), a([[0o13]]), "rsc", "MSB"),
    }


def test_trellis_tables_and_encoder_match_reference():
    g = np.load(os.path.join(GOLD, "trellis_tables.npz"))
    for name, (mem, gm, fb, ct, pf) in _specs().items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tr = Trellis(mem, gm, fb, ct, pf)
        assert np.array_equal(tr.next_state_table, g[name + "_next"]), name
        assert np.array_equal(tr.output_table, g[name + "_out"]), name
        msg = g[name + "_msg"]
        assert np.array_equal(conv_encode(msg, tr, "cont"), g[name + "_enc_cont"]), name
        assert np.array_equal(conv_encode(msg, tr, "term"), g[name + "_enc_term"]), name


def test_reference_kat_tables():
    """commpy/channelcoding/tests/test_convcode.py:23-127 (tables typed in from the reference's test)."""
    trs = helpers.reference_test_trellises()
    assert np.array_equal(trs[0].next_state_table, [[0, 2], [0, 2], [1, 3], [1, 3]])
    assert np.array_equal(trs[0].output_table, [[0, 3], [3, 0], [1, 2], [2, 1]])
    assert np.array_equal(trs[1].next_state_table, [[0, 2], [2, 0], [1, 3], [3, 1]])
    assert np.array_equal(trs[1].output_table, [[0, 3], [0, 3], [1, 2], [1, 2]])
    assert np.array_equal(trs[4].next_state_table, [[0, 1, 1, 0], [2, 3, 3, 2], [3, 2, 2, 3], [1, 0, 0, 1]])
    assert np.array_equal(trs[4].output_table, [[0, 3, 4, 7], [1, 2, 5, 6], [0, 3, 4, 7], [1, 2, 5, 6]])
    mes = np.array((0, 0, 1, 0))
    want = [[0, 0, 0, 0, 1, 1, 0, 1], [0, 0, 0, 0, 1, 1, 0, 1], [0, 0, 0, 1, 1, 0], [0, 0, 0, 1, 1, 0], [0, 0, 0, 1, 0, 0]]
    for tr, w in zip(trs, want):
        assert np.array_equal(conv_encode(mes, tr, "cont"), w)


def test_wifi_decimal_quirk_differs_from_octal():
    assert not np.array_equal(helpers.k7().output_table, helpers.k7_wifi_quirk().output_table)
    with pytest.raises(ValueError):
        Trellis(np.array([2]), np.array([[5, 7]]), polynomial_format="bogus")


def test_utilities():
    assert np.array_equal(dec2bitarray(17, 8), (0, 0, 0, 1, 0, 0, 0, 1))           # test_utilities.py:12
    assert np.array_equal(dec2bitarray((17, 12), 5), (1, 0, 0, 0, 1, 0, 1, 1, 0, 0))   # :13
    assert bitarray2dec(np.array([1, 0, 1, 1])) == 11
    assert hamming_dist(np.array([1, 0, 1]), np.array([0, 0, 1])) == 1
    assert euclid_dist(np.array([1.0, 2.0]), np.array([0.0, 0.0])) == 5.0
    assert np.array_equal(dec2bitarray(133, 7), (0, 0, 0, 0, 1, 0, 1))             # index-wrap quirk (utilities.py:81-85)


def test_puncturing_roundtrip_shapes():
    msg = np.arange(1, 21)
    pv = np.array([1, 1, 0, 1, 1, 0])
    p = puncturing(msg, pv)
    assert np.array_equal(p, msg[pv[np.arange(20) % 6] == 1])
    d = depuncturing(p.astype(float), pv, 20)
    assert np.array_equal(d[pv[np.arange(20) % 6] == 1], p) and (d[pv[np.arange(20) % 6] == 0] == 0).all()
    with pytest.raises(IndexError):
        depuncturing(p[:3], pv, 20)


def test_puncturing_depuncturing_and_punctured_encoder_match_reference_golden():
    """convcode.py:752-804 and :475-558 replayed from tests/golden/puncture.npz (oracle/make_puncture_golden.py ran the
    reference): same outputs, and the same exception type where the reference raises."""
    g = np.load(os.path.join(GOLD, "puncture.npz"))
    raised = 0
    for i in range(int(g["n_punct"])):
        k = "p%02d" % i
        msg, pv = g[k + "_msg"], g[k + "_pv"]
        p = puncturing(msg, pv)
        assert np.array_equal(p, g[k + "_punct"]), k
        err = str(g[k + "_err"])
        soft = p.astype(float) * 2 - 1
        if err:
            raised += 1
            with pytest.raises(IndexError):
                depuncturing(soft, pv, len(msg))
            assert err == "IndexError"
        else:
            assert np.array_equal(depuncturing(soft, pv, len(msg)), g[k + "_depunct"]), k
    tr = helpers.k7()
    for i in range(int(g["n_enc"])):
        k = "e%02d" % i
        got = conv_encode(g[k + "_msg"], tr, str(g[k + "_term"]), g[k + "_pm"])
        assert np.array_equal(got, g[k + "_coded"]), k


def test_interleaver_and_turbo_encode():
    il = RandInterlv(64, 1)
    assert np.array_equal(np.sort(il.p_array), np.arange(64))
    x = np.arange(64)
    assert np.array_equal(il.deinterlv(il.interlv(x)), x)
    from numpy.random import mtrand
    assert np.array_equal(il.p_array, mtrand.RandomState(1).permutation(np.arange(64)))
    tr = helpers.rsc_k4()
    msg = np.random.RandomState(0).randint(0, 2, 64)
    s, p1, p2 = turbo_encode(msg, tr, tr, il)
    assert len(s) == 64 and len(p1) == 64 and len(p2) == 131
    assert np.array_equal(s, msg)


def test_modem_constellations_match_reference():
    g = np.load(os.path.join(GOLD, "demod.npz"))
    custom = [re + im * 1j for re, im in product((-3.5, -0.5, 0.5, 3.5), repeat=2)]
    mods = {"psk4": PSKModem(4), "psk8": PSKModem(8), "psk16": PSKModem(16), "qam4": QAMModem(4),
            "qam16": QAMModem(16), "qam64": QAMModem(64), "qam256": QAMModem(256), "custom16": Modem(custom)}
    for name, md in mods.items():
        assert np.array_equal(np.asarray(md.constellation, dtype=np.complex128), g[name + "_constellation"]), name
    assert np.allclose(mods["qam256"].Es, 170) and np.allclose(mods["psk8"].Es, 1) and np.allclose(mods["custom16"].Es, 12.5)
    assert mods["qam64"].num_bits_symbol == 6 and mods["qam64"].m == 64
    with pytest.raises(ValueError):
        mods["qam16"].constellation = (0, 0, 0)
    with pytest.raises(ValueError):
        QAMModem(32)
    with pytest.raises(ValueError):
        PSKModem(6)
    # modulate: index = bits MSB first
    q = mods["qam16"]
    bits = np.array([1, 0, 1, 1, 0, 0, 0, 1])
    assert np.array_equal(q.modulate(bits), q.constellation[[11, 1]])
    assert signal_power(q.constellation) == q.Es


def test_ldpc_design_loader(tmp_path):
    from commpy_b200.channelcoding import write_ldpc_params
    rs = np.random.RandomState(5)
    H = (rs.rand(12, 24) < 0.3).astype(int)
    H[:, 0] = 1
    H[0, :] = 1
    path = os.path.join(tmp_path, "m.txt")
    write_ldpc_params(H, path)
    p = get_ldpc_code_params(path)
    assert p["n_vnodes"] == 24 and p["n_cnodes"] == 12
    from commpy_b200.channelcoding.ldpc import build_matrix
    import scipy.sparse as sp
    n_c = p["n_cnodes"]
    deg = p["cnode_deg_list"]
    adj = p["cnode_adj_list"].reshape(n_c, p["max_cnode_deg"])
    rows = np.repeat(np.arange(n_c), deg)
    cols = np.concatenate([adj[i, :deg[i]] for i in range(n_c)])
    H2 = sp.csc_matrix((np.ones(len(rows), np.int8), (rows, cols)), shape=(12, 24)).toarray()
    assert np.array_equal(H2, H)


def test_ldpc_design_loader_and_encoder_match_reference_golden(tmp_path):
    """ldpc.py:51-141 and :302-354 replayed from tests/golden/ldpc_design.npz (oracle/make_ldpc_design_golden.py ran the
    reference on its own design files): the design is re-written with write_ldpc_params from the stored matrix, loaded
    back, and every array of the parameter dict, the generator matrix and the encoded words must equal the reference's."""
    import scipy.sparse as sp
    from commpy_b200.channelcoding import triang_ldpc_systematic_encode, write_ldpc_params
    g = np.load(os.path.join(GOLD, "ldpc_design.npz"))
    for tag in ("g96", "w1440"):
        n, m = int(g[tag + "_n_vnodes"]), int(g[tag + "_n_cnodes"])
        H = sp.csc_matrix((np.ones(len(g[tag + "_H_rows"]), np.int8), (g[tag + "_H_rows"], g[tag + "_H_cols"])), shape=(m, n))
        path = os.path.join(tmp_path, tag + ".txt")
        write_ldpc_params(H.toarray(), path)
        p = get_ldpc_code_params(path, True)
        for k in ("n_vnodes", "n_cnodes", "max_cnode_deg", "max_vnode_deg", "cnode_deg_list", "vnode_deg_list"):
            assert np.array_equal(np.asarray(p[k]), g[tag + "_" + k]), (tag, k)
        # neighbour lists: the same sets per node (the Gallager file lists neighbours unsorted, the writer sorts them);
        # for the WiMax design, whose file is sorted, every array including the edge maps is identical
        for adj, deg, cnt in (("cnode_adj_list", "cnode_deg_list", m), ("vnode_adj_list", "vnode_deg_list", n)):
            a = np.asarray(p[adj]).reshape(cnt, -1)
            b = g[tag + "_" + adj].reshape(cnt, -1)
            for i in range(cnt):
                d = int(p[deg][i])
                assert sorted(a[i, :d]) == sorted(b[i, :d]), (tag, adj, i)
        if tag == "w1440":
            for k in ("cnode_adj_list", "cnode_vnode_map", "vnode_adj_list", "vnode_cnode_map"):
                assert np.array_equal(np.asarray(p[k]), g[tag + "_" + k]), (tag, k)
        assert (p["parity_check_matrix"] != H).nnz == 0
        assert np.array_equal(p["generator_matrix"].toarray(), g[tag + "_G"])
        assert np.array_equal(triang_ldpc_systematic_encode(g[tag + "_msg"], p, False), g[tag + "_coded"])
        assert np.array_equal(triang_ldpc_systematic_encode(g[tag + "_msg"][:, 0].copy(), p, False), g[tag + "_coded_padless_1d"])


def test_batch_encoder_helper_matches_conv_encode():
    rs = np.random.RandomState(9)
    for tr in helpers.reference_test_trellises() + [helpers.k7(), helpers.rsc_k4()]:
        for term in ("cont", "term"):
            msgs = rs.randint(0, 2, (5, 24 * tr.k))
            got = helpers.encode_batch(msgs, tr, term)
            for b in range(5):
                assert np.array_equal(got[b], conv_encode(msgs[b], tr, term)), (tr.k, tr.n, term)
