"""Modems: constellation construction / mapping on the host, soft and hard demapping on the GPU.

Mirror of commpy/modulation.py:39-262 (Modem, PSKModem, QAMModem).  `demodulate` runs in CUDA
(commpy_b200/csrc/demap.cu) through `cpb_demod_soft` / `cpb_demod_hard`; there is no CPU demapper.
OFDM helpers and the MIMO tree-search detectors of the reference are outside the decoding hot path.
"""
import ctypes as C

import numpy as np

from . import _lib
from .utilities import signal_power

__all__ = ["Modem", "PSKModem", "QAMModem"]


class _ModemBox:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr:
                _lib.load().cpb_modem_destroy(self.ptr)
        except Exception:
            pass


class Modem:
    """Custom modem (modulation.py:39-172).

    Attributes as in the reference: `constellation` (settable; length must be a power of two), `Es`, `m`,
    `num_bits_symbol`.  Symbol index k <-> bit pattern of k, MSB first.  With `reorder_as_gray=True` the given
    points are re-indexed so that `constellation[k] = given[gray^-1(k)]` (:68-77).
    """

    def __init__(self, constellation, reorder_as_gray=True):
        if reorder_as_gray:
            size = len(constellation)
            gray = np.arange(size) ^ (np.arange(size) >> 1)          # binary-reflected Gray sequence
            self.constellation = np.array(constellation)[gray.argsort()]
        else:
            self.constellation = constellation

    @property
    def constellation(self):
        return self._constellation

    @constellation.setter
    def constellation(self, value):
        num_bits_symbol = np.log2(len(value))
        if num_bits_symbol != int(num_bits_symbol):
            raise ValueError("Constellation length must be a power of 2.")
        self._constellation = np.array(value)
        self.Es = signal_power(self.constellation)
        self.m = self._constellation.size
        self.num_bits_symbol = int(num_bits_symbol)
        self._handles = {}

    def modulate(self, input_bits):
        """Map bits (MSB first, `num_bits_symbol` per symbol) to constellation points (modulation.py:79-98)."""
        bits = np.asarray(input_bits).astype(np.int64)
        nb = self.num_bits_symbol
        full = (len(bits) // nb) * nb
        idx = bits[:full].reshape(-1, nb) @ (1 << np.arange(nb - 1, -1, -1))
        if full != len(bits):                                   # a short last group is read as a shorter number
            tail = bits[full:]
            idx = np.append(idx, int(tail @ (1 << np.arange(len(tail) - 1, -1, -1))))
        return self._constellation[idx]

    def _handle(self):
        torch = _lib.require_cuda()
        dev = torch.cuda.current_device()
        box = self._handles.get(dev)
        if box is None:
            pts = np.ascontiguousarray(self._constellation, dtype=np.complex128)
            h = C.c_void_p()
            rc = _lib.load().cpb_modem_create(_lib.ptr(pts.view(np.float64)), int(self.m), C.byref(h))
            _lib.check(rc, "Modem constellation")
            box = _ModemBox(h)
            self._handles[dev] = box
        return box.ptr

    def demodulate_batch(self, input_symbols, demod_type, noise_var=0):
        """GPU demapper on a torch CUDA complex64 tensor (or numpy complex array) of any shape.

        Returns a torch CUDA tensor shaped input.shape + (num_bits_symbol,) flattened on the last two axes:
        float32 LLRs ('soft') or uint8 bits ('hard')."""
        torch = _lib.require_cuda()
        if demod_type not in ("hard", "soft"):
            raise ValueError('demod_type must be "hard" or "soft"')
        if hasattr(input_symbols, "data_ptr"):
            y = input_symbols if input_symbols.is_cuda else input_symbols.cuda()
            if y.dtype != torch.complex64:
                y = y.to(torch.complex64)
            y = y.contiguous()
        else:
            y = torch.from_numpy(np.ascontiguousarray(np.atleast_1d(input_symbols), dtype=np.complex64)).cuda()
        nsym = y.numel()
        nb = self.num_bits_symbol
        yr = torch.view_as_real(y)
        lib = _lib.load()
        if demod_type == "soft":
            if not noise_var > 0:
                raise ValueError("noise_var must be positive for soft demodulation")
            out = torch.empty((nsym * nb,), dtype=torch.float32, device=y.device)
            rc = lib.cpb_demod_soft(self._handle(), _lib.ptr(yr), C.c_int64(nsym), C.c_float(noise_var), _lib.ptr(out),
                                    _lib.stream_ptr(torch))
        else:
            out = torch.empty((nsym * nb,), dtype=torch.uint8, device=y.device)
            rc = lib.cpb_demod_hard(self._handle(), _lib.ptr(yr), C.c_int64(nsym), _lib.ptr(out), _lib.stream_ptr(torch))
        _lib.check(rc, "demodulate")
        return out.reshape(tuple(y.shape[:-1]) + (y.shape[-1] * nb,)) if y.dim() > 1 else out

    def demodulate_soft_host(self, input_symbols, noise_var):
        """Soft demapper for HOST arrays through the pipelined host entry point (cpb_demod_soft_host: chunked H2D / kernel /
        D2H on the modem handle's streams).  numpy complex array (any shape) -> float32 numpy array of LLRs, MSB first."""
        _lib.require_cuda()
        if not noise_var > 0:
            raise ValueError("noise_var must be positive for soft demodulation")
        y = np.ascontiguousarray(np.atleast_1d(input_symbols), dtype=np.complex64)
        out = np.empty(y.size * self.num_bits_symbol, dtype=np.float32)
        rc = _lib.load().cpb_demod_soft_host(self._handle(), _lib.ptr(y.view(np.float32)), C.c_int64(y.size),
                                             C.c_float(noise_var), _lib.ptr(out))
        _lib.check(rc, "demodulate")
        return out.reshape(tuple(y.shape[:-1]) + (y.shape[-1] * self.num_bits_symbol,)) if y.ndim > 1 else out

    def demodulate(self, input_symbols, demod_type, noise_var=0):
        """Drop-in for Modem.demodulate (modulation.py:100-141).

        'hard': nearest constellation point -> its bits (int8).  'soft': exact log-sum-exp LLRs
        log(sum_{bit=1} exp(-|y-c|^2/noise_var) / sum_{bit=0} ...), float64 array, MSB first per symbol.
        Computed in float32 on the GPU (relative error ~1e-4); finite where the reference under/overflows."""
        if demod_type == "soft":
            return self.demodulate_soft_host(np.atleast_1d(np.asarray(input_symbols)).reshape(-1), noise_var).astype(np.float64)
        out = self.demodulate_batch(np.atleast_1d(np.asarray(input_symbols)).reshape(-1), demod_type, noise_var)
        return out.cpu().numpy().astype(np.int8)


class PSKModem(Modem):
    """M-PSK: points exp(j*2*pi*i/m), Gray-labelled (modulation.py:175-210)."""

    def __init__(self, m):
        num_bits_symbol = np.log2(m)
        if num_bits_symbol != int(num_bits_symbol):
            raise ValueError("Constellation length must be a power of 2.")
        super().__init__(np.exp(1j * np.arange(0, 2 * np.pi, 2 * np.pi / m)))


class QAMModem(Modem):
    """Square M-QAM on the odd-integer grid, Gray-labelled per axis, no power normalisation
    (modulation.py:213-262): Es = 2(m-1)/3."""

    def __init__(self, m):
        side = np.sqrt(m)
        if side != int(side):
            raise ValueError("m must lead to a square QAM.")
        side = int(side)
        pam = np.arange(-side + 1, side, 2)
        # column-by-column snake through the grid: real part fixed per column, imaginary part up then down
        imag = np.tile(np.hstack((pam, pam[::-1])), side // 2)
        real = pam.repeat(side)
        super().__init__(imag * 1j + real)
