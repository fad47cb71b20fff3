"""Short run of the non-Viterbi kernels for ncu (LDPC at the DVB-S2 shape, 256-QAM demapper, MAP / turbo)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import helpers
from commpy_b200.channelcoding import RandInterlv, ldpc_bp_decode_batch, turbo_decode_batch
from commpy_b200.modulation import QAMModem

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "ldpc"):
    H = helpers.dvbs2_like_H()
    params = {"n_vnodes": 64800, "n_cnodes": 32400, "parity_check_matrix": H.tocsc()}
    sigma = 1.0 / (2 * 0.5 * 10 ** (1.0 / 10)) ** 0.5
    llr = (2.0 * (1.0 + sigma * torch.randn(256, 64800, device="cuda")) / sigma ** 2).float()
    ldpc_bp_decode_batch(llr, params, 4, "fp32", return_llrs=False)
if which in ("all", "demap"):
    q = QAMModem(256)
    y = torch.view_as_complex((torch.randn(1 << 24, 2, device="cuda") * 9).contiguous())
    for _ in range(3):
        q.demodulate_batch(y, "soft", 12.0)
if which in ("all", "turbo"):
    tr = helpers.rsc_k4()
    N, batch = 6144, 8192
    il = RandInterlv(N, 1)
    s2 = 1.0 / (2 * (1 / 3) * 10 ** (1.0 / 10))
    ys, y1, y2 = ((-1 + s2 ** 0.5 * torch.randn(batch, N, device="cuda")).float() for _ in range(3))
    turbo_decode_batch(ys, y1, y2, tr, s2, 1, il)
if which in ("all", "tx"):
    from commpy_b200.links import conv_link_tx
    for _ in range(2):
        conv_link_tx(helpers.k7(), QAMModem(256), 8192, 4096, 4, 0, 0.5)
torch.cuda.synchronize()
print("done")
