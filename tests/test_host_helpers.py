"""CPU checks of small host-side helpers of the package (no GPU, no library call)."""
import numpy as np
import pytest
import torch

from commpy_b200.channelcoding.turbo import _checked_perm, _host_out, suggest_map_window
from commpy_b200.links import kept_bits


def test_suggest_map_window_fills_the_gpu_and_stays_in_range():
    assert suggest_map_window(8192, 6144) == 1024           # 6 windows of 1024 steps: 49,152 threads
    assert suggest_map_window(1024, 6144) == 128            # 48 windows
    assert suggest_map_window(2048, 6144) == 256
    assert suggest_map_window(1, 6144) == 128               # never below 128 steps
    assert suggest_map_window(10 ** 6, 6144) == 1024        # never above the default
    for batch in (3, 77, 500, 4096):
        w = suggest_map_window(batch, 6144)
        assert w % 8 == 0 and 128 <= w <= 1024


def test_host_out_accepts_matching_buffers_only():
    a = _host_out(None, (3, 4), np.uint8, "out")
    assert a.shape == (3, 4) and a.dtype == np.uint8
    t = torch.empty((3, 4), dtype=torch.uint8)
    b = _host_out(t, (3, 4), np.uint8, "out")
    assert b.__array_interface__["data"][0] == t.data_ptr()            # the caller's memory, not a copy
    for bad in (np.empty((3, 5), np.uint8), np.empty((3, 4), np.int8), np.empty((4, 3), np.uint8).T, [[0] * 4] * 3):
        with pytest.raises(ValueError):
            _host_out(bad, (3, 4), np.uint8, "out")


def test_checked_perm_rejects_non_permutations_like_the_reference_would():
    class IL:
        pass
    il = IL()
    il.p_array = np.array([2, 0, 1, 3])
    assert _checked_perm(il, 4).dtype == np.int32
    il.p_array = np.array([2, 0, 1, 4])
    with pytest.raises(IndexError):                                     # out of range: the reference raises IndexError
        _checked_perm(il, 4)
    il.p_array = np.array([2, 0, 1, 1])
    with pytest.raises(ValueError):
        _checked_perm(il, 4)
    il.p_array = np.array([0, 1, 2])
    with pytest.raises(ValueError):
        _checked_perm(il, 4)


def test_kept_bits_counts_like_puncturing():
    from commpy_b200.channelcoding import puncturing
    for pv in ([1, 1, 1, 0], [1, 1, 1, 0, 0, 1], [1, 1, 1, 0, 0, 1, 1, 0, 0, 1]):
        for n in (0, 5, 24, 601):
            assert kept_bits(n, pv) == len(puncturing(np.zeros(n, dtype=int), pv))
    assert kept_bits(17, None) == 17
