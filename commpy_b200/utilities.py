"""Bit / distance helpers with CommPy's conventions (commpy/utilities.py:30-154).

Bit arrays are MSB first.  `dec2bitarray` keeps the reference's index wrap for numbers that need
more than `bit_width` bits (utilities.py:78-85): bit `p >= bit_width` lands on position
`2*bit_width - p - 1` of the MSB-first array, which is how `Wifi80211`'s DECIMAL (133, 171)
generator silently becomes taps (5, 43)."""
import numpy as np

__all__ = ["dec2bitarray", "decimal2bitarray", "bitarray2dec", "hamming_dist", "euclid_dist", "signal_power"]


def decimal2bitarray(number, bit_width):
    """MSB-first int8 bit array of one non-negative integer (utilities.py:58-86, wrap quirk included)."""
    number = int(number)
    bits = np.zeros(bit_width, np.int8)
    pos = 0
    while (1 << pos) <= number:
        if (number >> pos) & 1:
            idx = bit_width - pos - 1          # may be negative: Python-style wrap, IndexError past -bit_width
            if idx < -bit_width:
                raise IndexError("index %d is out of bounds for axis 0 with size %d" % (idx, bit_width))
            bits[idx] = 1
        pos += 1
    return bits


def dec2bitarray(in_number, bit_width):
    """Scalar or iterable of non-negative ints -> concatenated MSB-first bit arrays (utilities.py:30-55)."""
    if isinstance(in_number, (np.integer, int)):
        return decimal2bitarray(in_number, bit_width)
    nums = np.asarray(in_number)
    if nums.size and bit_width > 0 and nums.max(initial=0) < (1 << min(bit_width, 62)):
        shifts = np.arange(bit_width - 1, -1, -1)
        return ((nums.astype(np.int64)[:, None] >> shifts) & 1).astype(np.int8).reshape(-1)
    out = np.zeros(bit_width * len(nums), np.int8)
    for i, v in enumerate(nums):
        out[i * bit_width:(i + 1) * bit_width] = decimal2bitarray(v, bit_width)
    return out


def bitarray2dec(in_bitarray):
    """MSB-first bits -> integer (utilities.py:89-109)."""
    number = 0
    for b in in_bitarray:
        number = number * 2 + int(b)
    return number


def hamming_dist(in_bitarray_1, in_bitarray_2):
    """Number of differing positions of two 0/1 integer arrays (utilities.py:112-132)."""
    return np.bitwise_xor(in_bitarray_1, in_bitarray_2).sum()


def euclid_dist(in_array1, in_array2):
    """Squared euclidean distance (utilities.py:135-154)."""
    d = np.asarray(in_array1) - np.asarray(in_array2)
    return (d * d).sum()


def signal_power(signal):
    """Mean of |s|^2 (utilities.py:187-205)."""
    s = np.asarray(signal)
    return np.mean(np.abs(s) ** 2)
