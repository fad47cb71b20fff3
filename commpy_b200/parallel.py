"""Multi-GPU plumbing for the decoding path: independent frames shard across ranks with no data-path
collective; the only exchange is an all-reduce of the int64 error counters, which also
lets every rank take the same `bit_err < err_min and bit_send < send_max` decision (commpy/links.py:313).

One process per GPU, `torch.distributed` (NCCL on GPUs; gloo in the CPU tests)."""
import os


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process when unset)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced split of `n_items` frames: rank r owns [lo, hi).  Every frame is owned exactly once
    and the split does not depend on anything but (n_items, world_size), so results are reproducible."""
    if world_size < 1 or not (0 <= rank < world_size) or n_items < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def frame_seed(seed, frame_index):
    """Seed of one frame's random stream: a function of the GLOBAL frame index only, so the synthetic data (and
    therefore every error counter) is identical for 1, 2, 4 or 8 ranks."""
    return (int(seed) * 0x9E3779B97F4A7C15 + int(frame_index) * 0xBF58476D1CE4E5B9) % (1 << 63)


def batch_first_frame(batch_index, frames_per_batch, rank, world_size):
    """GLOBAL index of the first frame of batch `batch_index` on rank `rank` when every rank generates
    `frames_per_batch` frames per step: step b covers the frames [b*W*F, (b+1)*W*F), rank r the r-th slice of it.
    The frame streams are keyed by this index (cpb_conv_link_tx), so a BER point sees the same frames for any W."""
    if world_size < 1 or not (0 <= rank < world_size) or batch_index < 0 or frames_per_batch < 0:
        raise ValueError("bad batch request")
    return (int(batch_index) * int(world_size) + int(rank)) * int(frames_per_batch)


def allreduce_counters(counters, group=None):
    """Sum an int64 counter tensor over all ranks (no-op for a single process).  Returns the tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def stop_rule(counters, send_max, err_min):
    """links.py:313 evaluated on GLOBAL counters [bit_errors, frame_errors, bits_sent]: True = keep sending."""
    return int(counters[2]) < send_max and int(counters[0]) < err_min
