"""Short run of the soft-decision K=7 kernel for ncu (65,536 frames of N=1024)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import helpers
from commpy_b200.channelcoding import viterbi_decode_batch
tr = helpers.k7()
rs = np.random.RandomState(0)
_, x = helpers.channel_frames(tr, rs, 1024, 1024, "soft", "cont", flip=0.03, ebn0_db=4.0)
xt = torch.from_numpy(x.astype(np.float32)).cuda().repeat(64, 1).contiguous()
out = torch.empty((65536, 1024), dtype=torch.uint8, device="cuda")
for _ in range(2):
    viterbi_decode_batch(xt, tr, None, "soft", out=out)
torch.cuda.synchronize()
print("done")
