"""CPU, world_size 2, gloo: the sharding / counter all-reduce logic of the N>1 path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from commpy_b200 import parallel


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 64, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            seen = np.zeros(n, dtype=int)
            for r in range(w):
                lo, hi = parallel.shard_range(n, r, w)
                assert 0 <= lo <= hi <= n
                seen[lo:hi] += 1
            assert (seen == 1).all()
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)
    assert parallel.frame_seed(1, 5) != parallel.frame_seed(1, 6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(n_frames, rank, world)
    # every frame's "errors" are a function of its GLOBAL index only
    errs = np.array([parallel.frame_seed(3, f) % 5 for f in range(lo, hi)], dtype=np.int64)
    c = torch.tensor([int(errs.sum()), int((errs > 0).sum()), (hi - lo) * 1024], dtype=torch.int64)
    parallel.allreduce_counters(c)
    keep = parallel.stop_rule(c, send_max=10 ** 9, err_min=10)
    q.put((rank, c.tolist(), keep))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_counters_allreduce_world2_matches_single_process():
    n_frames = 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    errs = np.array([parallel.frame_seed(3, f) % 5 for f in range(n_frames)], dtype=np.int64)
    want = [int(errs.sum()), int((errs > 0).sum()), n_frames * 1024]
    for rank, c, keep in res:
        assert c == want, (rank, c, want)          # identical global counters on every rank, independent of world size
        assert keep is False


def test_batch_first_frame_tiles_the_frame_axis(monkeypatch):
    """Each step of W ranks x F frames covers a contiguous block of global frame indices exactly once, so the frames of
    a BER point (keyed by global index in cpb_conv_link_tx) do not depend on the number of ranks."""
    from commpy_b200 import parallel
    F = 7
    for W in (1, 2, 4, 8):
        seen = []
        for b in range(3):
            for r in range(W):
                first = parallel.batch_first_frame(b, F, r, W)
                seen.extend(range(first, first + F))
        assert seen == list(range(3 * W * F))
    # the same 8 x 7 frames, generated as one step of 8 ranks or as four steps of 2 ranks
    a = sorted(f for r in range(8) for f in range(parallel.batch_first_frame(0, F, r, 8), parallel.batch_first_frame(0, F, r, 8) + F))
    b = sorted(f for s_ in range(4) for r in range(2)
               for f in range(parallel.batch_first_frame(s_, F, r, 2), parallel.batch_first_frame(s_, F, r, 2) + F))
    assert a == b
    # ConvLinkGPU.make_batch asks the TX kernel for exactly that index (no GPU needed: the kernel call is stubbed)
    import commpy_b200.links as links
    from commpy_b200.modulation import QAMModem
    import helpers
    calls = []
    monkeypatch.setattr(links, "conv_link_tx", lambda tr, modem, frames, bits, seed, first, sigma, puncture=None: (calls.append((frames, bits, seed, first)) or (None, None)))
    monkeypatch.setenv("RANK", "3"); monkeypatch.setenv("WORLD_SIZE", "4")
    link = links.ConvLinkGPU(helpers.k7(), QAMModem(4), frame_bits=64, frames_per_batch=10, seed=9)
    link.make_batch(6.0, 5)
    assert calls == [(10, 64, 9, (5 * 4 + 3) * 10)]
