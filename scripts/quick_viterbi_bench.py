"""Scratch timing of the Viterbi kernels (device-resident inputs, CUDA events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import helpers
from commpy_b200.channelcoding import viterbi_decode_batch

tr = helpers.k7()
rs = np.random.RandomState(0)
for mode, nbits, batch in (("hard", 1024, 65536), ("soft", 1024, 65536), ("soft", 4096, 65536), ("unquantized", 1024, 65536)):
    _, x = helpers.channel_frames(tr, rs, 256, nbits, mode, "cont", flip=0.03, ebn0_db=4.0)
    xt = torch.from_numpy(x.astype(np.uint8 if mode == "hard" else np.float32)).cuda()
    xt = xt.repeat(batch // 256, 1).contiguous()
    out = torch.empty((batch, nbits), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        viterbi_decode_batch(xt, tr, None, mode, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        viterbi_decode_batch(xt, tr, None, mode, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s N=%d batch=%d: %.3f ms  %.3e cw/s" % (mode, nbits, batch, ms, batch / ms * 1e3), flush=True)
