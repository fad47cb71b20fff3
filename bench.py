#!/usr/bin/env python
"""bench.py -- codewords/sec of the K=7 rate-1/2 Viterbi hot path (N=1024) on N B200s, with roofline,
end-to-end (host buffers) and CPU-baseline figures on the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic frames that is already resident in HBM:
decode (cpb_viterbi_decode) + error count (cpb_count_errors) [+ NCCL all-reduce of the two int64 counters
when N > 1].  Frames shard across ranks with no data-path collective (weak scaling: 65,536 frames per GPU).
`e2e` times the same work through the public host-buffer API (pinned host memory, H2D and D2H inside the
timed region).  `--impl reference` times the CPU restatement of the reference's algorithm (oracle/, fp64,
all host threads) on the same workload -- the Python reference itself cannot travel to the GPU box
(BASELINE.md section 2 has its measured rate: ~1.5 codewords/s/core).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "codewords/sec (K=7 rate-1/2 Viterbi, N=1024)"
FALLBACK_HBM_GBS = 6650.0


# ------------------------------------------------------------------------------------------------ workloads
class ViterbiWorkload:
    """K=7 (0o133, 0o171) rate-1/2 'cont' frames through BSC (hard) or BPSK-AWGN (soft, LLR = 2y/sigma^2)."""

    def __init__(self, name, nbits, batch, mode, flip=0.03, ebn0_db=4.0):
        self.name, self.nbits, self.batch, self.mode = name, nbits, batch, mode
        self.flip, self.ebn0_db = flip, ebn0_db
        self.n_in = 2 * nbits
        self.dtype = "u8" if mode == "hard" else "int32 fixed-point metrics (f32 LLR in)"
        # coded values in + decoded bits out, u8 bits / f32 soft values
        self.alg_bytes = self.n_in * (1 if mode == "hard" else 4) + nbits
        self.kernel = "viterbi_fast_kernel_%s<FFCode<6,0133,0171>>" % ("hard" if mode == "hard" else "soft")

    def describe(self):
        d = {"workload": self.name, "code": "K=7 (0o133,0o171) rate 1/2, 'cont'", "info_bits": self.nbits,
             "coded_values_per_frame": self.n_in, "decoding_type": self.mode, "tb_depth": 30,
             "frames_per_gpu": self.batch, "alg_bytes_per_codeword": self.alg_bytes}
        d["channel"] = ("BSC p=%.3g" % self.flip) if self.mode == "hard" else ("BPSK AWGN Eb/N0=%.1f dB" % self.ebn0_db)
        return d

    def make(self, torch, seed, nbuf):
        """nbuf device-resident batches (distinct data), plus the transmitted messages."""
        import helpers
        from commpy_b200.channelcoding import conv_encode
        self.trellis = helpers.k7()
        g = torch.Generator(device="cuda")
        g.manual_seed(seed)
        self.inputs, self.msgs = [], []
        taps = [[b for b in range(7) if (poly >> b) & 1] for poly in (0o133, 0o171)]
        for _ in range(nbuf):
            msg = torch.randint(0, 2, (self.batch, self.nbits), generator=g, device="cuda", dtype=torch.uint8)
            pad = torch.nn.functional.pad(msg, (6, 0))
            coded = torch.empty((self.batch, self.n_in), dtype=torch.uint8, device="cuda")
            for j, tp in enumerate(taps):
                acc = torch.zeros_like(msg)
                for b in tp:
                    acc ^= pad[:, 6 - b:6 - b + self.nbits]
                coded[:, j::2] = acc
            if not hasattr(self, "_checked"):
                ref = conv_encode(msg[0].cpu().numpy(), self.trellis, "cont")
                assert np.array_equal(ref, coded[0].cpu().numpy()), "device encoder disagrees with conv_encode"
                self._checked = True
            if self.mode == "hard":
                flips = (torch.rand(coded.shape, generator=g, device="cuda") < self.flip).to(torch.uint8)
                x = coded ^ flips
            else:
                sigma2 = 1.0 / (2.0 * 0.5 * 10 ** (self.ebn0_db / 10.0))
                y = (2.0 * coded.float() - 1.0) + (sigma2 ** 0.5) * torch.randn(coded.shape, generator=g, device="cuda")
                x = (2.0 / sigma2) * y if self.mode == "soft" else y
            self.inputs.append(x.contiguous())
            self.msgs.append(msg)
        self.out = torch.empty((self.batch, self.nbits), dtype=torch.uint8, device="cuda")
        self.counters = torch.zeros(2, dtype=torch.int64, device="cuda")     # this rank's running totals
        self.glob = torch.zeros(2, dtype=torch.int64, device="cuda")         # all-reduced copy

    def decode(self, i):
        from commpy_b200.channelcoding import viterbi_decode_batch
        viterbi_decode_batch(self.inputs[i % len(self.inputs)], self.trellis, None, self.mode, out=self.out)

    def count(self, i, torch):
        import ctypes as C
        from commpy_b200 import _lib
        m = self.msgs[i % len(self.msgs)]
        rc = _lib.load().cpb_count_errors(_lib.ptr(self.out), _lib.ptr(m), C.c_int64(self.batch), C.c_int64(self.nbits),
                                          C.c_int64(self.nbits), C.c_int64(self.nbits), _lib.ptr(self.counters),
                                          _lib.stream_ptr(torch))
        _lib.check(rc, "count_errors")

    def host_buffers(self, torch):
        self.h_in = torch.empty(self.inputs[0].shape, dtype=self.inputs[0].dtype).pin_memory()
        self.h_in.copy_(self.inputs[0])
        self.h_out = torch.empty((self.batch, self.nbits), dtype=torch.uint8).pin_memory()
        return self.h_in.numel() * self.h_in.element_size(), self.h_out.numel()

    def e2e_step(self):
        from commpy_b200.channelcoding import viterbi_decode_batch
        viterbi_decode_batch(self.h_in, self.trellis, None, self.mode, out=self.h_out)

    def parity(self, torch, frames=48):
        """decode a few frames with the CPU oracle and compare (outside every timed region)"""
        from oracle import oracle
        self.decode(0)
        torch.cuda.synchronize()
        x = self.inputs[0][:frames].cpu().numpy().astype(np.float64)
        want = oracle.viterbi_decode_batch(x, self.trellis, None, self.mode, threads=min(8, os.cpu_count() or 1))
        got = self.out[:frames].cpu().numpy()
        return {"frames_checked": frames, "bit_mismatches_vs_oracle": int((got != want).sum()),
                "oracle_bit_errors": int((want != self.msgs[0][:frames].cpu().numpy()).sum()),
                "gpu_bit_errors": int((got != self.msgs[0][:frames].cpu().numpy()).sum())}

    def cpu_frames(self, frames, seed=0):
        """host-generated frames of the same recipe (float64, what the oracle eats)"""
        import helpers
        rs = np.random.RandomState(seed)
        _, x = helpers.channel_frames(helpers.k7(), rs, frames, self.nbits, self.mode, "cont", flip=self.flip,
                                      ebn0_db=self.ebn0_db)
        return x

    def cpu_run(self, x, threads):
        import helpers
        from oracle import oracle
        return oracle.viterbi_decode_batch(x, helpers.k7(), None, self.mode, threads=threads)


WORKLOADS = {
    "viterbi_k7_n1024_hard": lambda: ViterbiWorkload("viterbi_k7_n1024_hard", 1024, 65536, "hard"),
    "viterbi_k7_n1024_soft": lambda: ViterbiWorkload("viterbi_k7_n1024_soft", 1024, 65536, "soft"),
    "viterbi_k7_n4096_soft_c2": lambda: ViterbiWorkload("viterbi_k7_n4096_soft_c2", 4096, 65536, "soft"),
}
DEFAULT_WORKLOAD = "viterbi_k7_n1024_hard"


# ------------------------------------------------------------------------------------------------ helpers
def usable_cores():
    """host threads this process may really use: CPU affinity, capped by the cgroup CPU quota when there is one"""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return max(1, n)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy bandwidth)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML on a host thread DURING the timed region."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.reasons = [], set()
        self.mx = None
        self._stop = False
        self._t = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            while not self._stop:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                    r = int(fn(h))
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
                time.sleep(0.004)
        except Exception as e:          # NVML missing: report no samples rather than fail the bench
            self.err = repr(e)

    def start(self):
        import threading
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        time.sleep(0.05)

    def stop(self):
        self._stop = True
        if self._t is not None:
            self._t.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [], "samples": 0, "error": getattr(self, "err", None)}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def timed_cpu(wl, threads, target_s):
    """oracle on a bounded sample: pilot to size it, then one timed run"""
    pilot = wl.cpu_frames(max(8, 2 * threads), seed=1)
    t0 = time.perf_counter()
    wl.cpu_run(pilot, threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    frames = int(min(wl.batch, max(len(pilot), len(pilot) * target_s / dt)))
    x = wl.cpu_frames(frames, seed=2)
    t0 = time.perf_counter()
    wl.cpu_run(x, threads)
    dt = time.perf_counter() - t0
    return frames / dt, frames, dt


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]()
    threads = usable_cores()
    per_step = []
    frames_step = None
    for s in range(args.warmup + args.steps):
        rate, frames, dt = timed_cpu(wl, threads, target_s=2.0)
        frames_step = frames
        if s >= args.warmup:
            per_step.append((frames, dt))
    tot_f = sum(f for f, _ in per_step)
    tot_t = sum(t for _, t in per_step)
    value = tot_f / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "codewords/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": wl.describe(),
        "cpu_baseline": {"value": value, "unit": "codewords/s", "cores": threads, "kind": "port",
                         "sample": "%d frames per step of the same recipe, oracle/commpy_oracle.c (fp64 restatement of "
                                   "convcode.py:561-749) over %d host threads" % (frames_step, threads)},
        "e2e": {"value": value, "unit": "codewords/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the Python reference cannot run on the GPU box; its measured rate in the build container is "
                "~1.5 codewords/s/core (BASELINE.md section 2)",
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from commpy_b200 import parallel
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    wl = WORKLOADS[args.workload]()
    wl.make(torch, seed=1000 + rank, nbuf=3)        # 3 distinct batches: the working set is > 3x the 126 MB L2
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, ev=None):
        if ev is not None:
            ev[0].record()
        wl.decode(i)
        if ev is not None:
            ev[1].record()
        wl.count(i, torch)
        wl.glob.copy_(wl.counters)
        parallel.allreduce_counters(wl.glob)        # the only collective: two int64 error counters (no-op at N=1)

    for i in range(W):
        step(i)
    barrier()
    wl.counters.zero_()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0.record()
    for i in range(K):
        step(W + i, kev[i])
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([ms_total, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = float(t[0]), float(t[1])
    counters = wl.glob.cpu().numpy().astype(np.int64)
    value = n_gpus * wl.batch * K / (ms_total / 1e3)
    bits_total = n_gpus * wl.batch * K * wl.nbits

    # ---- end to end through the public host-buffer API (pinned host memory in, pinned host memory out)
    h2d, d2h = wl.host_buffers(torch)
    Ke = max(3, min(K, 10))
    wl.e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        wl.e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_gpus * wl.batch * Ke / float(te[0])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    achieved = wl.alg_bytes * wl.batch / (kernel_ms / 1e3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "codewords/s", "n_gpus": n_gpus, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": wl.dtype, "data": "synthetic",
        "config": dict(wl.describe(), parallelism="frames sharded over %d GPU(s), no data-path collective" % n_gpus,
                       l2="inputs cycle over 3 distinct device batches (%.0f MB each, > 126 MB L2 in total)"
                          % (wl.inputs[0].numel() * wl.inputs[0].element_size() / 1e6)),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(wl.kernel), "kernel": wl.kernel, "kernel_ms": kernel_ms,
                     "peak_source": peak_src,
                     "note": "ACS-issue bound, not HBM bound: 65,856 add-compare-selects per codeword (DESIGN.md)"},
        "e2e": {"value": e2e_value, "unit": "codewords/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": Ke, "api": "commpy_b200.channelcoding.viterbi_decode_batch(pinned host array) -> cpb_viterbi_decode_host"},
        "gpu_launches": 2 * K,
        "clocks": clocks,
        "ber": {"bit_errors": int(counters[0]), "frame_errors": int(counters[1]), "bits": int(bits_total),
                "ber": float(counters[0]) / bits_total},
    }
    if n_gpus == 1:
        line["parity"] = wl.parity(torch)
        threads = usable_cores()
        rate, frames, dt = timed_cpu(wl, threads, target_s=12.0)
        line["cpu_baseline"] = {"value": rate, "unit": "codewords/s", "cores": threads, "kind": "port",
                                "sample": "%d frames of the same recipe in %.1f s: oracle/commpy_oracle.c (fp64 restatement "
                                          "of convcode.py:561-749), %d host threads" % (frames, dt, threads)}
        if not args.no_extras:
            line["extras"] = run_extras(torch)
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_extras(torch):
    """Other rows of the hot path, timed briefly on the same GPU (device-resident inputs, CUDA events)."""
    out = {}
    peak, _ = measured_peak()

    def timeit(fn, reps, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    for name in ("viterbi_k7_n1024_soft", "viterbi_k7_n4096_soft_c2"):
        try:
            wl = WORKLOADS[name]()
            wl.make(torch, seed=7, nbuf=2)
            ms = timeit(lambda: wl.decode(0), reps=5, warm=2)
            out[name] = {"value": wl.batch / ms * 1e3, "unit": "codewords/s", "ms": ms, "frames": wl.batch,
                         "roofline_frac": wl.alg_bytes * wl.batch / ms / 1e6 / peak, "parity": wl.parity(torch, 16)}
            del wl
            torch.cuda.empty_cache()
        except Exception as e:          # an extra must never take the headline down
            out[name] = {"error": repr(e)[:200]}
    try:
        import helpers
        from commpy_b200.channelcoding import RandInterlv, turbo_decode_batch
        tr = helpers.rsc_k4()
        N, batch = 6144, 8192          # config C3's batch
        il = RandInterlv(N, 1)
        s2 = 1.0 / (2 * (1 / 3) * 10 ** (1.0 / 10))
        ys, y1, y2 = ((-1 + s2 ** 0.5 * torch.randn(batch, N, device="cuda")).float() for _ in range(3))
        ms = timeit(lambda: turbo_decode_batch(ys, y1, y2, tr, s2, 6, il), reps=2, warm=1)
        out["turbo_k4_n6144_6it_c3"] = {"value": batch / ms * 1e3, "unit": "codewords/s", "ms": ms, "frames": batch,
                                        "roofline_frac": 79872.0 * batch / ms / 1e6 / peak}
    except Exception as e:
        out["turbo_k4_n6144_6it_c3"] = {"error": repr(e)[:200]}
    try:
        import helpers
        from commpy_b200.channelcoding import ldpc_bp_decode_batch
        H = helpers.dvbs2_like_H()
        params = {"n_vnodes": 64800, "n_cnodes": 32400, "parity_check_matrix": H.tocsc()}
        batch, iters = 256, 50
        sigma = 1.0 / (2 * 0.5 * 10 ** (1.0 / 10)) ** 0.5           # Eb/N0 = 1 dB: this graph never converges -> all 50 iterations run
        llr = (2.0 * (1.0 + sigma * torch.randn(batch, 64800, device="cuda")) / sigma ** 2).float()
        work = llr.clone()
        def run_ldpc():
            work.copy_(llr)
            return ldpc_bp_decode_batch(work, params, iters, "fp32", return_llrs=False, return_iters=True)
        ms = timeit(lambda: run_ldpc(), reps=2, warm=1)
        _, it = run_ldpc()
        mean_it = float(it.float().mean())
        bytes_fi = 12 * H.nnz + 8 * 64800
        out["ldpc_dvbs2shape_64800_minsum_c4"] = {
            "value": batch / ms * 1e3, "unit": "codewords/s", "ms": ms, "frames": batch, "mean_iterations": mean_it,
            "matrix": "DVB-S2-SHAPED surrogate (tests/helpers.py::dvbs2_like_H), 226,799 edges",
            "roofline_frac": (batch * (mean_it * bytes_fi + 5 * 64800)) / ms / 1e6 / peak}
        del llr, work
        torch.cuda.empty_cache()
    except Exception as e:
        out["ldpc_dvbs2shape_64800_minsum_c4"] = {"error": repr(e)[:200]}
    try:
        import helpers
        from commpy_b200.links import ConvLinkGPU
        from commpy_b200.modulation import QAMModem
        link = ConvLinkGPU(helpers.k7(), QAMModem(256), frame_bits=4096, frames_per_batch=49152, decoding_type="soft", seed=4)
        snr = 14.0 + 10 * np.log10(8)                                # Eb/N0 = 14 dB (SNR = Eb/N0 + 10 log10(bits/symbol))
        msg, y, nv = link.make_batch(snr, 0, torch)
        cnt = torch.zeros(3, dtype=torch.int64, device="cuda")
        ms = timeit(lambda: link.receive_decode_count(msg, y, nv, cnt, torch), reps=3, warm=1)
        nsym = y.numel()
        out["c5_rx_chain_qam256_k7_soft"] = {
            "value": nsym / ms * 1e3, "unit": "symbols/s", "ms": ms, "symbols": nsym,
            "chain": "cpb_demod_soft -> cpb_viterbi_decode(soft) -> cpb_count_errors, symbols resident in HBM, 49,152 frames of 4096 bits",
            "roofline_frac": 12.0 * nsym / ms / 1e6 / peak}
        # the whole BER point of config 5 through the public API: TX kernel (cpb_conv_link_tx) + RX chain + stop rule,
        # 2 batches of 49,152 frames = 1.0e8 symbols at Eb/N0 = 14 dB
        ms_tx = timeit(lambda: link.make_batch(snr, 1, torch), reps=3, warm=1)
        link.link_performance([snr], send_max=1, err_min=10 ** 12)       # warm-up: one batch (allocator, handles)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        send_max = 2 * 49152 * 4096
        bers = link.link_performance([snr], send_max=send_max - 1, err_min=10 ** 12)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["c5_link_performance_1e8_symbols"] = {
            "value": 2 * nsym / dt, "unit": "symbols/s", "seconds": dt, "symbols": 2 * nsym, "ber": float(bers[0]),
            "tx_ms_per_batch": ms_tx,
            "chain": "ConvLinkGPU.link_performance: cpb_conv_link_tx -> cpb_demod_soft -> cpb_viterbi_decode(soft) -> "
                     "cpb_count_errors -> counter all-reduce, nothing leaves the device but 3 counters per batch"}
        del msg, y
        torch.cuda.empty_cache()
    except Exception as e:
        out["c5_rx_chain_qam256_k7_soft"] = {"error": repr(e)[:200]}
    try:
        from commpy_b200.modulation import QAMModem
        q = QAMModem(256)
        n = 1 << 24
        y = torch.view_as_complex((torch.randn(n, 2, device="cuda") * 9).contiguous())
        ms = timeit(lambda: q.demodulate_batch(y, "soft", 12.0), reps=5, warm=2)
        out["demap_qam256_soft"] = {"value": n / ms * 1e3, "unit": "symbols/s", "ms": ms,
                                    "roofline_frac": 40.0 * n / ms / 1e6 / peak}
    except Exception as e:
        out["demap_qam256_soft"] = {"error": repr(e)[:200]}
    return out


_JSON_FD = None


def emit(line):
    """The one JSON line of the contract, written to the REAL stdout (see main())."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # stdout carries exactly one JSON line.  Libraries may print to fd 1 (NCCL's version banner does when NCCL_DEBUG is
    # set in the environment), so fd 1 is pointed at stderr for the whole run and the line goes to a saved copy.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
