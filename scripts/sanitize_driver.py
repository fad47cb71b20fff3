"""Small invocation of every kernel family (used under compute-sanitizer by scripts/sanitize.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import helpers
from commpy_b200.channelcoding import (RandInterlv, ldpc_bp_decode_batch, map_decode_batch, turbo_decode_batch, viterbi_decode_batch,
                                        viterbi_decode_punctured_batch, puncturing)
from commpy_b200.links import ConvLinkGPU
from commpy_b200.modulation import QAMModem, PSKModem

rs = np.random.RandomState(0)
tr = helpers.k7()
# Viterbi: hard (byte and packed), soft, unquantized, punctured, generic trellis
_, x = helpers.channel_frames(tr, rs, 70, 200, "hard", "cont", flip=0.06)
viterbi_decode_batch(x.astype(np.uint8), tr, None, "hard")
_, x = helpers.channel_frames(tr, rs, 70, 256, "hard", "cont", flip=0.06)
viterbi_decode_batch(np.packbits(x.astype(np.uint8), axis=1), tr, None, "hard", packed=True)
for mode in ("soft", "unquantized"):
    _, x = helpers.channel_frames(tr, rs, 40, 130, mode, "term", ebn0_db=2.0)
    viterbi_decode_batch(x.astype(np.float32), tr, 15, mode)
pv = [1, 1, 1, 0, 0, 1]
_, x = helpers.channel_frames(tr, rs, 40, 300, "soft", "cont", ebn0_db=3.0)
viterbi_decode_punctured_batch(np.stack([puncturing(r, pv) for r in x]).astype(np.float32), tr, pv, x.shape[1])
g = helpers.reference_test_trellises()[2]
_, x = helpers.channel_frames(g, rs, 20, 120 * g.k, "hard", "cont", flip=0.05)
viterbi_decode_batch(x.astype(np.uint8), g, None, "hard")
# BCJR / turbo
rsc = helpers.rsc_k4()
N = 2048
ys, y1, y2 = (torch.randn(6, N, device="cuda") * 0.8 - 1 for _ in range(3))
map_decode_batch(ys, y1, rsc, 0.64, torch.zeros(6, N, device="cuda"))
turbo_decode_batch(ys, y1, y2, rsc, 0.64, 2, RandInterlv(N, 1))
turbo_decode_batch(ys[:, :516].contiguous(), y1[:, :516].contiguous(), y2[:, :516].contiguous(), rsc, 0.64, 2, RandInterlv(516, 2),
                   torch.randn(6, 516, device="cuda"))            # step-major loop, 4-step tail segment, a-priori L_int
ys, y1 = (torch.randn(5, 301, device="cuda") - 1 for _ in range(2))
map_decode_batch(ys, y1, rsc, 0.7, torch.zeros(5, 301, device="cuda"))
# LDPC: bulk-copy check pass (>= 128 frames), small batch, fp64, SPA
import scipy.sparse as sp
gl = np.load(os.path.join(ROOT, "tests", "golden", "ldpc.npz"))
rel, nblk, iters, m, n = gl["l03_meta"]
H = sp.csr_matrix((np.ones(len(gl["l03_indices"]), np.int8), gl["l03_indices"], gl["l03_indptr"]), shape=(int(m), int(n)))
params = {"n_vnodes": int(n), "n_cnodes": int(m), "parity_check_matrix": H.tocsc()}
sigma = 0.8
llr = (2.0 * (1.0 + sigma * rs.randn(160, int(n))) / sigma ** 2)
ldpc_bp_decode_batch(llr.astype(np.float32), params, 6, "fp32")
ldpc_bp_decode_batch(llr[:9].astype(np.float32), params, 6, "fp32")
ldpc_bp_decode_batch(llr[:9].copy(), params, 4, "fp64")
ldpc_bp_decode_batch(llr[:9].astype(np.float32), params, 4, "fp32", decoder_algorithm="SPA")
# demapper (separable, general, hard) and the TX + link chain
y = torch.view_as_complex(torch.randn(5000, 2, device="cuda") * 3)
QAMModem(64).demodulate_batch(y, "soft", 1.5)
PSKModem(8).demodulate_batch(y, "soft", 0.5)
QAMModem(16).demodulate_batch(y, "hard")
link = ConvLinkGPU(tr, QAMModem(16), frame_bits=512, frames_per_batch=64, decoding_type="soft", seed=1)
link.link_performance([9.0], send_max=100000, err_min=10 ** 9)
link = ConvLinkGPU(helpers.k7_wifi_quirk(), QAMModem(256), frame_bits=600, frames_per_batch=64, decoding_type="soft", seed=1, puncture=pv)
link.link_performance([27.0], send_max=100000, err_min=10 ** 9)
torch.cuda.synchronize()
print("sanitize driver ok")
