"""Scratch timing of the LDPC / turbo / demap kernels (device-resident inputs, CUDA events)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import helpers
from commpy_b200.channelcoding import RandInterlv, ldpc_bp_decode_batch, turbo_decode_batch
from commpy_b200.modulation import QAMModem


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# demap 256-QAM
q = QAMModem(256)
n = 1 << 24
y = torch.view_as_complex((torch.randn(n, 2, device="cuda") * 9).contiguous())
ms = timeit(lambda: q.demodulate_batch(y, "soft", 12.0), reps=5, warm=2)
print("demap 256-QAM %d symbols: %.3f ms  %.3e sym/s  %.1f GB/s (40 B/sym)" % (n, ms, n / ms * 1e3, n * 40 / ms / 1e6), flush=True)

# turbo C3: N=6144, 6 it
tr = helpers.rsc_k4()
N, batch = 6144, 1024
il = RandInterlv(N, 1)
s2 = 1.0 / (2 * (1 / 3) * 10 ** (1.0 / 10))
ys = (-1 + np.sqrt(s2) * torch.randn(batch, N, device="cuda")).float()
y1 = (-1 + np.sqrt(s2) * torch.randn(batch, N, device="cuda")).float()
y2 = (-1 + np.sqrt(s2) * torch.randn(batch, N, device="cuda")).float()
ms = timeit(lambda: turbo_decode_batch(ys, y1, y2, tr, s2, 6, il), reps=2, warm=1)
print("turbo N=6144 6it batch %d: %.2f ms  %.3e cw/s" % (batch, ms, batch / ms * 1e3), flush=True)

# LDPC: DVB-S2-shaped surrogate is built in bench.py; here WiMax 1440 with a large batch
import scipy.sparse as sp
g = np.load(os.path.join(ROOT, "tests", "golden", "ldpc.npz"))
rel, nblk, iters, m, nn = g["l03_meta"]
H = sp.csr_matrix((np.ones(len(g["l03_indices"]), np.int8), g["l03_indices"], g["l03_indptr"]), shape=(int(m), int(nn)))
params = {"n_vnodes": int(nn), "parity_check_matrix": H.tocsc()}
batch = 16384
sigma = 1.0 / np.sqrt(2 * 0.5 * 10 ** (1.0 / 10))
llr = (2.0 * (1.0 + sigma * torch.randn(batch, int(nn), device="cuda")) / sigma ** 2).float()
E = H.nnz
ms = timeit(lambda: ldpc_bp_decode_batch(llr.clone(), params, 50, "fp32", return_llrs=False), reps=2, warm=1)
dec, it = ldpc_bp_decode_batch(llr.clone(), params, 50, "fp32", return_llrs=False, return_iters=True)
mean_it = float(it.float().mean())
print("ldpc wimax1440 batch %d 50it: %.2f ms  %.3e cw/s  mean iters %.1f  ~%.0f GB/s (12E+8n per frame-iter)" % (
    batch, ms, batch / ms * 1e3, mean_it, batch * mean_it * (12 * E + 8 * int(nn)) / ms / 1e6), flush=True)
