"""Interleavers (commpy/channelcoding/interleavers.py:13-77)."""
import numpy as np
from numpy.random import mtrand

__all__ = ["RandInterlv"]


class _Interleaver:
    def interlv(self, in_array):
        """out[i] = in[p_array[i]] (interleavers.py:13-29)."""
        return np.asarray(in_array)[self.p_array]

    def deinterlv(self, in_array):
        """out[p_array[i]] = in[i] (interleavers.py:31-47)."""
        in_array = np.asarray(in_array)
        out = np.zeros(len(in_array), in_array.dtype)
        out[self.p_array] = in_array
        return out


class RandInterlv(_Interleaver):
    """Random interleaver: `RandomState(seed).permutation(arange(length))` (interleavers.py:50-77)."""

    def __init__(self, length, seed):
        self.p_array = mtrand.RandomState(seed).permutation(np.arange(length))
