"""Scratch: ConvLinkGPU.link_performance sweep of config 5 at one GPU's share of an 8-GPU run (12,288 frames per point) and at
the full single-GPU share (98,304), with and without issuing the next point before the last counters are read."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, helpers
from commpy_b200.links import ConvLinkGPU
from commpy_b200.modulation import QAMModem

snrs = [e + 10 * math.log10(8) for e in (8.0, 10.0, 12.0, 14.0, 16.0)]
for fpb in (12288, 98304):
    link = ConvLinkGPU(helpers.k7(), QAMModem(256), frame_bits=4096, frames_per_batch=fpb, decoding_type="soft", seed=4)
    for overlap in (False, True):
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            bers = link.link_performance(snrs, send_max=fpb * 4096, err_min=10 ** 12, stop_early=False, overlap_points=overlap)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("frames per point %6d overlap %d: %.3f ms per 5-point sweep, %.3e symbols/s  BERs %s" % (
            fpb, overlap, dt * 1e3, 5 * fpb * 1024 / dt, ["%.3e" % b for b in bers]), flush=True)
