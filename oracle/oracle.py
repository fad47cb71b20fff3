"""ctypes front-end of the CPU oracle (oracle/commpy_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of commpy_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The function signatures mirror the reference's so parity
tests read like the reference's own tests:

    viterbi_decode   <- commpy/channelcoding/convcode.py:661
    map_decode       <- commpy/channelcoding/turbo.py:163
    turbo_decode     <- commpy/channelcoding/turbo.py:254
    ldpc_bp_decode   <- commpy/channelcoding/ldpc.py:144   (MSA only)
    demodulate       <- commpy/modulation.py:100

Trellis / interleaver / modem arguments are duck-typed: anything exposing the
reference's attributes (k, n, total_memory, number_states, next_state_table,
output_table / p_array / constellation) works -- the reference's own objects
and commpy_b200's host-side mirrors alike.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcommpy_oracle.so")
_lib = None

_MODES = {"hard": 0, "soft": 1, "unquantized": 2}


def build(force=False):
    """Compile libcommpy_oracle.so with gcc (seconds)."""
    src = os.path.join(_HERE, "commpy_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-std=c11",
             "-shared", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        for name in ("orc_viterbi_decode", "orc_viterbi_decode_batch", "orc_map_decode", "orc_turbo_decode",
                     "orc_turbo_decode_batch", "orc_ldpc_minsum", "orc_ldpc_sumproduct", "orc_demod_soft", "orc_demod_hard"):
            getattr(_lib, name).restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _tables(trellis):
    nst = np.ascontiguousarray(trellis.next_state_table, dtype=np.int32)
    out = np.ascontiguousarray(trellis.output_table, dtype=np.int32)
    return nst, out


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("oracle %s failed with code %d" % (what, rc))


def viterbi_decode(coded_bits, trellis, tb_depth=None, decoding_type="hard"):
    if decoding_type not in _MODES:
        raise ValueError('The available decoding types are "hard", "soft" and "unquantized')
    coded = np.ascontiguousarray(coded_bits, dtype=np.float64)
    nst, out = _tables(trellis)
    k, n = int(trellis.k), int(trellis.n)
    L = int(len(coded) * (k / n))
    dec = np.zeros(L, dtype=np.int64)
    rc = lib().orc_viterbi_decode(_p(coded, C.c_double), C.c_int64(len(coded)), _p(nst, C.c_int32),
                                  _p(out, C.c_int32), k, n, int(trellis.total_memory),
                                  int(trellis.number_states), int(tb_depth or 0), _MODES[decoding_type],
                                  _p(dec, C.c_int64))
    _check(rc, "viterbi_decode")
    return dec


def viterbi_decode_batch(coded, trellis, tb_depth=None, decoding_type="hard", threads=1):
    """coded: (batch, len) array.  Returns (batch, L) int64.  threads>1 splits frames over host threads."""
    coded = np.ascontiguousarray(coded, dtype=np.float64)
    batch, ln = coded.shape
    nst, out = _tables(trellis)
    k, n = int(trellis.k), int(trellis.n)
    L = int(ln * (k / n))
    dec = np.zeros((batch, L), dtype=np.int64)
    fn = lib().orc_viterbi_decode_batch

    def run(lo, hi):
        if hi > lo:
            rc = fn(_p(coded[lo:hi], C.c_double), C.c_int64(hi - lo), C.c_int64(ln), _p(nst, C.c_int32),
                    _p(out, C.c_int32), k, n, int(trellis.total_memory), int(trellis.number_states),
                    int(tb_depth or 0), _MODES[decoding_type], _p(dec[lo:hi], C.c_int64), 1)
            _check(rc, "viterbi_decode_batch")

    _split(run, batch, threads)
    return dec


def _split(run, batch, threads):
    threads = max(1, min(int(threads), batch))
    if threads == 1:
        run(0, batch)
        return
    # ctypes releases the GIL during the foreign call, so plain threads scale across cores.
    chunk = -(-batch // (threads * 4))
    edges = [(lo, min(batch, lo + chunk)) for lo in range(0, batch, chunk)]
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda e: run(*e), edges))


def map_decode(sys_symbols, non_sys_symbols, trellis, noise_variance, L_int, mode="decode"):
    sys_ = np.ascontiguousarray(sys_symbols, dtype=np.float64)
    par = np.ascontiguousarray(non_sys_symbols, dtype=np.float64)
    La = np.ascontiguousarray(L_int, dtype=np.float64)
    N = len(sys_)
    nst, out = _tables(trellis)
    L_out = np.zeros(N)
    bits = np.zeros(N, dtype=np.int64)
    rc = lib().orc_map_decode(_p(sys_, C.c_double), _p(par, C.c_double), C.c_int64(N), _p(nst, C.c_int32),
                              _p(out, C.c_int32), int(trellis.number_states), int(trellis.number_inputs),
                              C.c_double(noise_variance), _p(La, C.c_double), 1 if mode == "decode" else 0,
                              _p(L_out, C.c_double), _p(bits, C.c_int64))
    _check(rc, "map_decode")
    return [L_out, bits]


def turbo_decode(sys_symbols, non_sys_symbols_1, non_sys_symbols_2, trellis, noise_variance,
                 number_iterations, interleaver, L_int=None):
    sys_ = np.ascontiguousarray(sys_symbols, dtype=np.float64)
    p1 = np.ascontiguousarray(non_sys_symbols_1, dtype=np.float64)
    p2 = np.ascontiguousarray(non_sys_symbols_2, dtype=np.float64)
    perm = np.ascontiguousarray(interleaver.p_array, dtype=np.int64)
    N = len(sys_)
    nst, out = _tables(trellis)
    bits = np.zeros(N, dtype=np.int64)
    La = None if L_int is None else np.ascontiguousarray(L_int, dtype=np.float64)
    rc = lib().orc_turbo_decode(_p(sys_, C.c_double), _p(p1, C.c_double), _p(p2, C.c_double), C.c_int64(N),
                                _p(nst, C.c_int32), _p(out, C.c_int32), int(trellis.number_states),
                                int(trellis.number_inputs), C.c_double(noise_variance), int(number_iterations),
                                _p(perm, C.c_int64), None if La is None else _p(La, C.c_double),
                                _p(bits, C.c_int64))
    _check(rc, "turbo_decode")
    return bits


def turbo_decode_batch(sys_, p1, p2, trellis, noise_variance, number_iterations, interleaver, threads=1):
    sys_ = np.ascontiguousarray(sys_, dtype=np.float64)
    p1 = np.ascontiguousarray(p1, dtype=np.float64)
    p2 = np.ascontiguousarray(p2, dtype=np.float64)
    perm = np.ascontiguousarray(interleaver.p_array, dtype=np.int64)
    batch, N = sys_.shape
    nst, out = _tables(trellis)
    bits = np.zeros((batch, N), dtype=np.int64)
    fn = lib().orc_turbo_decode_batch

    def run(lo, hi):
        if hi > lo:
            rc = fn(_p(sys_[lo:hi], C.c_double), _p(p1[lo:hi], C.c_double), _p(p2[lo:hi], C.c_double),
                    C.c_int64(hi - lo), C.c_int64(N), _p(nst, C.c_int32), _p(out, C.c_int32),
                    int(trellis.number_states), int(trellis.number_inputs), C.c_double(noise_variance),
                    int(number_iterations), _p(perm, C.c_int64), _p(bits[lo:hi], C.c_int64), 1)
            _check(rc, "turbo_decode_batch")

    _split(run, batch, threads)
    return bits


def csr_from_params(ldpc_code_params):
    """(row_ptr, col_idx) int32 CSR of params['parity_check_matrix'] (any scipy sparse / dense)."""
    import scipy.sparse as sp
    H = sp.csr_matrix(ldpc_code_params["parity_check_matrix"])
    H.sort_indices()
    return (np.ascontiguousarray(H.indptr, dtype=np.int32), np.ascontiguousarray(H.indices, dtype=np.int32),
            H.shape[0], H.shape[1])


def ldpc_bp_decode(llr_vec, ldpc_code_params, decoder_algorithm, n_iters, return_iters=False, threads=1):
    """'MSA' or 'SPA'.  llr_vec is clipped in place like the reference (ldpc.py:186) when it is a float64 array."""
    if decoder_algorithm not in ("MSA", "SPA"):
        raise NameError('Please input a valid decoder_algorithm string (meanning "SPA" or "MSA").')
    row_ptr, col_idx, m, n = csr_from_params(ldpc_code_params)
    if isinstance(llr_vec, np.ndarray) and llr_vec.dtype == np.float64 and llr_vec.flags.c_contiguous:
        llr = llr_vec
    else:
        llr = np.ascontiguousarray(llr_vec, dtype=np.float64).copy()
    n_blocks = llr.size // n
    dec = np.zeros(n_blocks * n, dtype=np.int8)
    out = np.zeros(n_blocks * n, dtype=np.float64)
    iters = np.zeros(n_blocks, dtype=np.int32)
    fn = lib().orc_ldpc_minsum if decoder_algorithm == "MSA" else lib().orc_ldpc_sumproduct
    flat = llr.reshape(-1)

    def run(lo, hi):
        if hi > lo:
            rc = fn(_p(flat[lo * n:hi * n], C.c_double), C.c_int64(hi - lo), n, m, _p(row_ptr, C.c_int32),
                    _p(col_idx, C.c_int32), int(n_iters), _p(dec[lo * n:hi * n], C.c_int8),
                    _p(out[lo * n:hi * n], C.c_double), _p(iters[lo:hi], C.c_int32))
            _check(rc, "ldpc_minsum")

    _split(run, n_blocks, threads)
    dec_word = dec.reshape(-1, n_blocks, order="F").squeeze().astype(np.int8)     # ldpc.py:251-254
    out_llrs = out.reshape(-1, n_blocks, order="F").squeeze()
    if return_iters:
        return dec_word, out_llrs, iters
    return dec_word, out_llrs


def demodulate(modem, input_symbols, demod_type, noise_var=0):
    y = np.ascontiguousarray(np.atleast_1d(input_symbols), dtype=np.complex128)
    cst = np.ascontiguousarray(modem.constellation, dtype=np.complex128)
    M = len(cst)
    nb = int(np.log2(M))
    if demod_type == "soft":
        out = np.zeros(len(y) * nb)
        rc = lib().orc_demod_soft(_p(y.view(np.float64), C.c_double), C.c_int64(len(y)),
                                  _p(cst.view(np.float64), C.c_double), M, C.c_double(noise_var),
                                  _p(out, C.c_double))
    elif demod_type == "hard":
        out = np.zeros(len(y) * nb, dtype=np.int8)
        rc = lib().orc_demod_hard(_p(y.view(np.float64), C.c_double), C.c_int64(len(y)),
                                  _p(cst.view(np.float64), C.c_double), M, _p(out, C.c_int8))
    else:
        raise ValueError('demod_type must be "hard" or "soft"')
    _check(rc, "demodulate")
    return out
