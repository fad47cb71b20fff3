"""Scratch: TX chain of config 5 (256-QAM, K=7, 4096-bit frames), word-parallel kernel against the bit-serial one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
from commpy_b200 import _lib
from commpy_b200.links import conv_link_tx
from commpy_b200.modulation import QAMModem
tr, modem, frames = helpers.k7(), QAMModem(256), 8192
for force in (1, 0):
    _lib.set_option(_lib.OPT_TX_FORCE_GENERIC, force)
    for _ in range(2): conv_link_tx(tr, modem, frames, 4096, 4, 0, 0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): conv_link_tx(tr, modem, frames, 4096, 4, 0, 0.5)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%s kernel: %.3f ms per %d symbols = %.3e symbols/s" % ("bit-serial" if force else "word-parallel", ms, frames * 1024, frames * 1024 / ms * 1e3), flush=True)
_lib.set_option(_lib.OPT_TX_FORCE_GENERIC, 0)
