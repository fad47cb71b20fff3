"""Scratch timing of cpb_turbo_decode at config-3 frame length for several batch sizes and MAP window lengths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, helpers
from commpy_b200 import _lib
from commpy_b200.channelcoding import RandInterlv, turbo_decode_batch
rsc = helpers.rsc_k4(); N = 6144; il = RandInterlv(N, 1)
s2 = 1.0 / (2 * (1 / 3) * 10 ** (1.0 / 10))
for batch in (8192, 4096, 2048, 1024):
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    y = [(-1 + s2 ** 0.5 * torch.randn(batch, N, device="cuda", generator=g)).float() for _ in range(3)]
    ref = None
    for win in (0, 512, 256, 128):
        _lib.set_option(_lib.OPT_BCJR_WINDOW, win)
        for _ in range(2):
            out = turbo_decode_batch(y[0], y[1], y[2], rsc, s2, 6, il)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = turbo_decode_batch(y[0], y[1], y[2], rsc, s2, 6, il)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        if ref is None: ref = out.clone()
        print("batch %5d window %4s: %.3f ms  %.3e cw/s  bits differing from the 1024-step windows: %d of %d" % (batch, win or 1024, ms, batch / ms * 1e3, int((out != ref).sum()), out.numel()), flush=True)
    _lib.set_option(_lib.OPT_BCJR_WINDOW, 0)
