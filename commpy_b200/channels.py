"""Flat-fading channels for the link models (host side): the callers' side of the decoding path.

Mirror of the parts of commpy/channels.py the link tests use: `SISOFlatChannel` (fixed complex gain + AWGN) and
`MIMOFlatChannel` (Kronecker-model flat fading, uncorrelated Rayleigh by default, + AWGN).  The random stream is consumed in
the reference's order (noise: real part then imaginary part, channels.py:49-55; gains: real then imaginary, :366-370), so a
seeded `numpy.random` run reproduces the reference's transmissions exactly.
"""
import numpy as np

__all__ = ["SISOFlatChannel", "MIMOFlatChannel"]


class _FlatChannel:
    def __init__(self):
        self.noises = None
        self.channel_gains = None
        self.unnoisy_output = None
        self.noise_std = None

    def generate_noises(self, dims):
        assert self.noise_std is not None, "Noise standard deviation must be set before propagation."
        if self.isComplex:                                        # channels.py:52-53: noise_std / 2 per real component
            re = np.random.standard_normal(dims)
            self.noises = (re + 1j * np.random.standard_normal(dims)) * self.noise_std * 0.5
        else:
            self.noises = np.random.standard_normal(dims) * self.noise_std

    def set_SNR_dB(self, SNR_dB, code_rate=1.0, Es=1):
        """channels.py:57-74: noise_std = sqrt((isComplex + 1) nb_tx Es / (code_rate 10^(SNR/10)))"""
        self.noise_std = np.sqrt((self.isComplex + 1) * self.nb_tx * Es / (code_rate * 10 ** (SNR_dB / 10)))

    def set_SNR_lin(self, SNR_lin, code_rate=1, Es=1):
        self.noise_std = np.sqrt((self.isComplex + 1) * self.nb_tx * Es / (code_rate * SNR_lin))

    @property
    def isComplex(self):
        return self._isComplex


class SISOFlatChannel(_FlatChannel):
    """y = gain * x + noise with gain = fading_param[0] + fading_param[1] * N(0,1)-type term (channels.py:100-240); the link
    tests use the pure AWGN case fading_param = (1 + 0j, 0j)."""

    def __init__(self, noise_std=None, fading_param=(1, 0)):
        super().__init__()
        self.noise_std = noise_std
        self.fading_param = fading_param

    @property
    def fading_param(self):
        return self._fading_param

    @fading_param.setter
    def fading_param(self, fading_param):
        if fading_param[1] + abs(fading_param[0]) ** 2 != 1:
            raise ValueError("With this parameters, the channel would add or remove energy.")
        self._fading_param = fading_param
        self._isComplex = isinstance(fading_param[0], complex)

    @property
    def nb_tx(self):
        return 1

    @property
    def nb_rx(self):
        return 1

    def propagate(self, msg):
        msg = np.asarray(msg)
        if np.iscomplexobj(msg) and not self.isComplex:
            raise TypeError("Trying to propagate a complex message in a real channel.")
        n = len(msg)
        self.generate_noises(n)
        if self.isComplex:
            re = np.random.standard_normal(n)
            g = (re + 1j * np.random.standard_normal(n)) * np.sqrt(0.5 * self.fading_param[1])
        else:
            g = np.random.standard_normal(n) * np.sqrt(self.fading_param[1])
        self.channel_gains = g + self.fading_param[0]
        self.unnoisy_output = self.channel_gains * msg
        return self.unnoisy_output + self.noises


def _sqrtm_psd(a):
    """principal square root of a (Hermitian positive semi-definite) correlation matrix"""
    a = np.asarray(a)
    w, v = np.linalg.eigh(a)
    return (v * np.sqrt(np.clip(w, 0, None))) @ v.conj().T


class MIMOFlatChannel(_FlatChannel):
    """nb_rx x nb_tx flat fading, one independent channel matrix per transmitted symbol vector (channels.py:242-384):
    H = Rr^(1/2) G Rt^(1/2)^T + mean with G i.i.d. unit-variance Gaussian; fading_param = (mean, Rt, Rr)."""

    def __init__(self, nb_tx, nb_rx, noise_std=None, fading_param=None):
        super().__init__()
        self.nb_tx, self.nb_rx = nb_tx, nb_rx
        self.noise_std = noise_std
        self.fading_param = fading_param if fading_param is not None else (np.zeros((nb_rx, nb_tx)), np.identity(nb_tx), np.identity(nb_rx))

    @property
    def fading_param(self):
        return self._fading_param

    @fading_param.setter
    def fading_param(self, fading_param):
        nlos = np.trace(np.kron(np.asarray(fading_param[1]).T, fading_param[2]))
        los = np.sum(np.abs(fading_param[0]) ** 2)
        if abs(nlos + los - self.nb_tx * self.nb_rx) > 1e-3:
            raise ValueError("With this parameters, the channel would add or remove energy.")
        self._fading_param = fading_param
        self._isComplex = np.iscomplexobj(fading_param[0])

    def uncorr_rayleigh_fading(self, dtype):
        """uncorrelated Rayleigh fading of the given dtype (channels.py:477-485)"""
        self.fading_param = np.zeros((self.nb_rx, self.nb_tx), dtype), np.identity(self.nb_tx), np.identity(self.nb_rx)

    def propagate(self, msg):
        msg = np.asarray(msg)
        if np.iscomplexobj(msg) and not self.isComplex:
            raise TypeError("Trying to propagate a complex message in a real channel.")
        nb_vect, rest = divmod(len(msg), self.nb_tx)
        if rest:                                                   # zero padding to whole symbol vectors
            msg = np.hstack((msg, np.zeros(self.nb_tx - rest)))
            nb_vect += 1
        msg = msg.reshape(nb_vect, -1)
        self.generate_noises((nb_vect, self.nb_rx))
        dims = (nb_vect, self.nb_rx, self.nb_tx)
        if self.isComplex:
            re = np.random.standard_normal(dims)
            g = (re + 1j * np.random.standard_normal(dims)) * np.sqrt(0.5)
        else:
            g = np.random.standard_normal(dims)
        mean, rt, rr = self.fading_param
        g = np.einsum("ij,ajk,lk->ail", _sqrtm_psd(rr), g, _sqrtm_psd(rt))
        self.channel_gains = g + mean
        self.unnoisy_output = np.einsum("ijk,ik->ij", self.channel_gains, msg)
        return self.unnoisy_output + self.noises
