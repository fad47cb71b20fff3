"""Monte-Carlo link simulation around the GPU decoding path (next row of SURVEY.md section 8f).

* `LinkModel` / `link_performance` keep the reference's sequential, callable-driven loop and signature
  (commpy/links.py:29-64, :269-342) -- host logic only, the callables it is given do the work (e.g. this
  package's `Modem.demodulate` and `viterbi_decode`).
* `AwgnSisoChannel` is the one channel convention the decoding path needs to synthesise its inputs
  (commpy/channels.py:53,74: `noise_std = sqrt((isComplex+1)*nb_tx*Es / (rate*10^(SNR/10)))`, complex noise
  `(N(0,1) + jN(0,1)) * noise_std * 0.5`).  Fading / MIMO channels are out of scope.
* `ConvLinkGPU` is the batched form for config C5: random bits -> convolutional encoder -> Modem.modulate ->
  AWGN generated on the device by one CUDA kernel (`conv_link_tx` -> cpb_conv_link_tx, counter-based Philox
  randomness keyed by the GLOBAL frame index), then this package's demapper, Viterbi decoder and error
  counter; frames shard over ranks and the error counters are all-reduced so every rank takes the same
  stop decision (`links.py:313`).
"""
import math
from fractions import Fraction
from inspect import getfullargspec

import numpy as np

from . import _lib, parallel

__all__ = ["link_performance", "LinkModel", "AwgnSisoChannel", "ConvLinkGPU", "conv_link_tx", "idd_decoder"]


class AwgnSisoChannel:
    """SISO AWGN channel with the reference's SNR convention (channels.py:37-93, :181-221 with fading (1+0j, 0))."""

    nb_tx = 1
    nb_rx = 1

    def __init__(self, is_complex=True, rng=None):
        self.isComplex = bool(is_complex)
        self.noise_std = None
        self.channel_gains = 1.0
        self.rng = rng if rng is not None else np.random

    def set_SNR_dB(self, SNR_dB, code_rate=1, Es=1):
        self.noise_std = math.sqrt((self.isComplex + 1) * self.nb_tx * Es / (code_rate * 10 ** (SNR_dB / 10)))

    def generate_noises(self, dims):
        if self.isComplex:
            return (self.rng.standard_normal(dims) + 1j * self.rng.standard_normal(dims)) * self.noise_std * 0.5
        return self.rng.standard_normal(dims) * self.noise_std

    def propagate(self, msg):
        msg = np.asarray(msg)
        self.noises = self.generate_noises(len(msg))
        # SISOFlatChannel draws its fading variates after the noise even when their variance is zero
        # (channels.py:213-217); the same draws are made and discarded here so that a seeded run consumes the random
        # stream exactly like the reference with fading_param = (1 + 0j, 0j) and reproduces its BERs to the last digit.
        self.rng.standard_normal(len(msg))
        if self.isComplex:
            self.rng.standard_normal(len(msg))
        self.channel_gains = np.ones(len(msg), dtype=complex if self.isComplex else float)
        self.unnoisy_output = msg
        return msg + self.noises


def link_performance(link_model, SNRs, send_max, err_min, send_chunk=None, code_rate=1):
    """Same as `link_model.link_performance(...)` (links.py:29-64)."""
    if not send_chunk:
        send_chunk = err_min
    return link_model.link_performance(SNRs, send_max, err_min, send_chunk, code_rate)


class LinkModel:
    """Link model built from callables, as in the reference (links.py:67-153):
    `modulate(bits) -> symbols`, `channel` (set_SNR_dB / propagate / channel_gains / noise_std / nb_tx),
    `receive(y, H, constellation, noise_var) -> bits or LLRs`, `decoder(array)` or the 6-argument form
    `decoder(y, H, constellation, noise_var, array, bits_per_send)` chosen by arity (links.py:306)."""

    def __init__(self, modulate, channel, receive, num_bits_symbol, constellation, Es=1, decoder=None, rate=Fraction(1, 1),
                 number_chunks_per_send=1, stop_on_surpass_error=True):
        self.modulate = modulate
        self.channel = channel
        self.receive = receive
        self.num_bits_symbol = num_bits_symbol
        self.constellation = constellation
        self.Es = Es
        self.rate = rate
        self.number_chunks_per_send = number_chunks_per_send
        self.stop_on_surpass_error = stop_on_surpass_error
        self.decoder = (lambda msg: msg) if decoder is None else decoder
        self.full_simulation_results = None

    def link_performance(self, SNRs, send_max, err_min, send_chunk=None, code_rate=1):
        """Sequential Monte-Carlo BER estimate, one chunk per iteration (links.py:269-342)."""
        BERs = np.zeros_like(SNRs, dtype=float)
        if send_chunk is None:
            send_chunk = err_min
        if type(code_rate) is float:
            code_rate = Fraction(code_rate).limit_denominator(100)
        self.rate = code_rate
        divider = (Fraction(1, self.num_bits_symbol * self.channel.nb_tx) * 1 / code_rate).denominator
        send_chunk = max(divider, send_chunk // divider * divider)
        receive_size = self.channel.nb_tx * self.num_bits_symbol
        full_args_decoder = len(getfullargspec(self.decoder).args) > 1
        for i, snr in enumerate(SNRs):
            self.channel.set_SNR_dB(snr, float(code_rate), self.Es)
            bit_send = 0
            bit_err = 0
            while bit_send < send_max and bit_err < err_min:
                msg = np.random.choice((0, 1), send_chunk)
                y = self.channel.propagate(self.modulate(msg))
                nv = self.channel.noise_std ** 2
                if np.ndim(y) > 1:           # one received vector per channel use (MIMO-shaped channels)
                    received = np.empty(int(math.ceil(len(msg) / float(self.rate))))
                    for j in range(len(y)):
                        received[receive_size * j:receive_size * (j + 1)] = \
                            self.receive(y[j], self.channel.channel_gains[j], self.constellation, nv)
                else:
                    received = self.receive(y, self.channel.channel_gains, self.constellation, nv)
                if full_args_decoder:
                    decoded = self.decoder(y, self.channel.channel_gains, self.constellation, nv, received,
                                           self.channel.nb_tx * self.num_bits_symbol)
                else:
                    decoded = self.decoder(received)
                bit_err += np.bitwise_xor(msg, np.asarray(decoded)[:len(msg)].astype(int)).sum()
                bit_send += send_chunk
            BERs[i] = bit_err / bit_send
            if bit_err < err_min:
                break
        return BERs


    def _one_transmission(self, msg, receive_size, full_args_decoder):
        y = self.channel.propagate(self.modulate(msg))
        nv = self.channel.noise_std ** 2
        if np.ndim(y) > 1:
            received = np.empty(int(math.ceil(len(msg) / float(self.rate))))
            for j in range(len(y)):
                received[receive_size * j:receive_size * (j + 1)] = \
                    self.receive(y[j], self.channel.channel_gains[j], self.constellation, nv)
        else:
            received = self.receive(y, self.channel.channel_gains, self.constellation, nv)
        if full_args_decoder:
            return self.decoder(y, self.channel.channel_gains, self.constellation, nv, received,
                                self.channel.nb_tx * self.num_bits_symbol)
        return self.decoder(received)

    def link_performance_full_metrics(self, SNRs, tx_max, err_min, send_chunk=None, code_rate=1,
                                      number_chunks_per_send=1, stop_on_surpass_error=True):
        """Per-transmission metrics (links.py:155-267): returns (BERs, BEs, CEs, NCs) and caches them on
        `full_simulation_results`.  A transmission carries `number_chunks_per_send` chunks encoded and decoded as ONE
        stream (:229-230, :250); errors are then counted per chunk (:252-256)."""
        BERs = np.zeros_like(SNRs, dtype=float)
        BEs = np.zeros((len(SNRs), tx_max), dtype=int)
        CEs = np.zeros((len(SNRs), tx_max), dtype=int)
        NCs = np.zeros((len(SNRs), tx_max), dtype=int)
        if send_chunk is None:
            send_chunk = err_min
        if type(code_rate) is float:
            code_rate = Fraction(code_rate).limit_denominator(100)
        self.rate = code_rate
        divider = (Fraction(1, self.num_bits_symbol * self.channel.nb_tx) * 1 / code_rate).denominator
        send_chunk = max(divider, send_chunk // divider * divider)
        receive_size = self.channel.nb_tx * self.num_bits_symbol
        full_args_decoder = len(getfullargspec(self.decoder).args) > 1
        for i, snr in enumerate(SNRs):
            self.channel.set_SNR_dB(snr, float(code_rate), self.Es)
            sent = 0
            bit_err = np.zeros(tx_max, dtype=int)
            chunk_count = np.zeros(tx_max, dtype=int)
            for tx in range(tx_max):
                if stop_on_surpass_error and bit_err.sum() > err_min:
                    break
                msg = np.random.choice((0, 1), send_chunk * number_chunks_per_send)
                decoded = np.asarray(self._one_transmission(msg, receive_size, full_args_decoder))
                for c in range(number_chunks_per_send):
                    sl = slice(send_chunk * c, send_chunk * (c + 1))
                    bit_err[tx] += np.bitwise_xor(msg[sl], decoded[sl].astype(int)).sum()
                chunk_count[tx] += number_chunks_per_send
                sent += 1
            BERs[i] = bit_err.sum() / (sent * send_chunk)
            BEs[i] = bit_err
            CEs[i] = np.where(bit_err > 0, 1, 0)
            NCs[i] = chunk_count
            if BEs[i].sum() < err_min:
                break
        self.full_simulation_results = BERs, BEs, CEs, NCs
        return BERs, BEs, CEs, NCs


def _ff_taps(trellis):
    """Generator taps (delay 0 = current input) of a k=1 feed-forward shift-register trellis, or None."""
    if trellis.k != 1:
        return None
    M, n = trellis.total_memory, trellis.n
    nst, otab = np.asarray(trellis.next_state_table), np.asarray(trellis.output_table)
    S = trellis.number_states
    for s in range(S):
        for u in range(2):
            if nst[s, u] != ((u << (M - 1)) | (s >> 1)):
                return None
    taps = np.zeros((n, M + 1), dtype=np.int64)
    for j in range(n):
        taps[j, 0] = (otab[0, 1] >> (n - 1 - j)) & 1
        for b in range(1, M + 1):
            taps[j, b] = (otab[1 << (M - b), 0] >> (n - 1 - j)) & 1
    for s in range(S):          # the code must be linear in (state, input) for the tap form to hold
        for u in range(2):
            regs = [u] + [(s >> (M - b)) & 1 for b in range(1, M + 1)]
            sym = 0
            for j in range(n):
                sym = (sym << 1) | (int(np.dot(taps[j], regs)) & 1)
            if sym != otab[s, u]:
                return None
    return taps


def conv_link_tx(trellis, modem, frames, frame_bits, seed, first_frame, noise_sigma, puncture=None):
    """Device-side TX chain of `frames` frames starting at GLOBAL frame index `first_frame`: random message ->
    conv_encode(..., 'cont') -> [puncturing(coded, puncture)] -> modem.modulate -> + noise_sigma * (N(0,1) + jN(0,1)).

    Returns (msg uint8 (frames, frame_bits), y complex64 (frames, frame_bits*n/bits_per_symbol)) as CUDA tensors.
    The streams depend only on (seed, global frame index): any split of the frames over calls or ranks gives the
    same frames."""
    import ctypes as C
    from .channelcoding.convcode import _trellis_handle
    torch = _lib.require_cuda()
    nb = int(modem.num_bits_symbol)
    kept = kept_bits(int(trellis.n) * int(frame_bits), puncture)
    if kept % nb:
        raise ValueError("the (punctured) coded bits of a frame must fill whole symbols")
    nsym = kept // nb
    msg = torch.empty((int(frames), int(frame_bits)), dtype=torch.uint8, device="cuda")
    y = torch.empty((int(frames), nsym), dtype=torch.complex64, device="cuda")
    if puncture is None:
        rc = _lib.load().cpb_conv_link_tx(_trellis_handle(trellis), modem._handle(), C.c_int64(int(frames)),
                                          C.c_int64(int(frame_bits)), C.c_uint64(int(seed) & ((1 << 64) - 1)),
                                          C.c_int64(int(first_frame)), C.c_float(float(noise_sigma)), _lib.ptr(msg),
                                          _lib.ptr(y), _lib.stream_ptr(torch))
    else:
        pv = np.ascontiguousarray(puncture, dtype=np.int32)
        rc = _lib.load().cpb_conv_link_tx_punctured(_trellis_handle(trellis), modem._handle(), C.c_int64(int(frames)),
                                                    C.c_int64(int(frame_bits)), C.c_uint64(int(seed) & ((1 << 64) - 1)),
                                                    C.c_int64(int(first_frame)), C.c_float(float(noise_sigma)),
                                                    _lib.ptr(pv), int(len(pv)), _lib.ptr(msg), _lib.ptr(y),
                                                    _lib.stream_ptr(torch))
    _lib.check(rc, "conv_link_tx")
    return msg, y


def kept_bits(n_coded, puncture):
    """how many of n_coded bits puncturing(message, puncture) keeps (convcode.py:752-774)"""
    if puncture is None:
        return int(n_coded)
    pv = np.asarray(puncture)
    return int(np.sum(pv[np.arange(int(n_coded)) % len(pv)] == 1))


class ConvLinkGPU:
    """Batched convolutional-code link over AWGN on the GPU(s): TX chain and RX chain in this package's CUDA.

    Parameters: `trellis` (k=1 feed-forward, e.g. the K=7 (0o133,0o171) code), `modem` (commpy_b200 Modem),
    `frame_bits` information bits per frame ('cont' termination), `frames_per_batch` frames decoded per step and rank.
    """

    def __init__(self, trellis, modem, frame_bits=4096, frames_per_batch=4096, decoding_type="soft", tb_depth=None, seed=0,
                 puncture=None):
        taps = _ff_taps(trellis)
        if taps is None:
            raise NotImplementedError("ConvLinkGPU generates frames on the device for k=1 feed-forward codes only")
        if decoding_type not in ("soft", "hard"):
            raise ValueError("decoding_type must be 'soft' or 'hard'")
        self.trellis, self.modem, self.taps = trellis, modem, taps
        self.frame_bits, self.frames = int(frame_bits), int(frames_per_batch)
        self.decoding_type, self.tb_depth, self.seed = decoding_type, tb_depth, int(seed)
        nb = modem.num_bits_symbol
        # puncturing pattern over the coded stream (802.11: [1,1,1,0] = 2/3, [1,1,1,0,0,1] = 3/4, ...): punctured on the
        # device in the TX kernel, depunctured inside the Viterbi kernel's load
        self.puncture = None if puncture is None else [int(v) for v in puncture]
        if self.puncture is not None and decoding_type != "soft":
            raise ValueError("a punctured link decodes soft values")
        n_coded = trellis.n * self.frame_bits
        kept = kept_bits(n_coded, self.puncture)
        if kept % nb:
            raise ValueError("the (punctured) coded bits of a frame must fill whole symbols")
        self.rate = Fraction(trellis.k, trellis.n) * Fraction(n_coded, kept)

    # -- TX chain on the device ---------------------------------------------------------------------
    def noise_std(self, snr_db):
        """channels.py:74: noise_std = sqrt(2 Es / (rate 10^(SNR/10))); each real component gets noise_std / 2."""
        return math.sqrt(2 * self.modem.Es / (float(self.rate) * 10 ** (snr_db / 10)))

    def make_batch(self, snr_db, batch_index, torch=None):
        """(msg bits, received symbols, noise_var) for one batch of this rank -- everything stays on the device.
        Batch `batch_index` of rank r covers the global frames [(batch_index*world + r) * frames, ... + frames)."""
        rank, world, _ = parallel.world()
        first = parallel.batch_first_frame(batch_index, self.frames, rank, max(world, 1))
        ns = self.noise_std(snr_db)
        msg, y = conv_link_tx(self.trellis, self.modem, self.frames, self.frame_bits, self.seed, first, 0.5 * ns,
                              self.puncture)
        return msg, y, ns ** 2                                                                     # links.py:329

    # -- RX chain: this package's kernels -------------------------------------------------------------
    def receive_decode_count(self, msg, y, noise_var, counters, torch):
        import ctypes as C
        from .channelcoding import viterbi_decode_batch
        if self.decoding_type == "soft":
            rx = self.modem.demodulate_batch(y, "soft", noise_var)
        else:
            rx = self.modem.demodulate_batch(y, "hard")
        if self.puncture is None:
            dec = viterbi_decode_batch(rx, self.trellis, self.tb_depth, self.decoding_type)
        else:
            from .channelcoding import viterbi_decode_punctured_batch
            dec = viterbi_decode_punctured_batch(rx, self.trellis, self.puncture, self.trellis.n * self.frame_bits,
                                                 self.tb_depth, "soft")
        L = msg.shape[1]
        rc = _lib.load().cpb_count_errors(_lib.ptr(dec), _lib.ptr(msg), C.c_int64(msg.shape[0]), C.c_int64(L),
                                          C.c_int64(dec.shape[1]), C.c_int64(L), _lib.ptr(counters), _lib.stream_ptr(torch))
        _lib.check(rc, "count_errors")
        return dec

    def link_performance(self, SNRs, send_max, err_min, return_counters=False, stop_early=True, overlap_points=True):
        """BER per SNR (dB, `SNR = Eb/N0 + 10 log10(bits/symbol)` as in the reference's examples,
        conv_encode_decode.py:102; the code rate enters through `set_SNR_dB`, channels.py:74).  A point ends
        when the GLOBAL counters reach `err_min` errors or `send_max` bits; like the reference the sweep stops after
        the first point that ends below `err_min` errors (links.py:339-341).  The stop rule is evaluated once per
        batch (frames_per_batch * world_size frames), not once per frame.

        Nothing on the critical path waits for the host: the bit count of a point is known in advance, and the error
        count of batch b is read (from a pinned snapshot) while batch b+1 is already running -- a batch issued after the
        error threshold was crossed is discarded, so the counters are exactly those of the in-order rule.
        With return_counters=True also returns the summed [bit errors, frame errors, bits] tensor of all points.
        stop_early=False measures every point of the sweep (the reference abandons the sweep after the first point that
        ends below err_min errors).
        The counters of a point's LAST batch are read after the next point's first batch has been issued (that batch is
        discarded if the sweep ends there), so the GPU does not idle between points either; overlap_points=False reads them
        at once -- same BERs and counters, the cross-check of the test suite."""
        torch = _lib.require_cuda()
        BERs = np.zeros(len(SNRs))
        batch_index = 0
        _, world, _ = parallel.world()
        bits_per_batch = self.frames * max(world, 1) * self.frame_bits
        grand = torch.zeros(3, dtype=torch.int64, device="cuda")
        pinned = torch.empty((8, 3), dtype=torch.int64).pin_memory()       # snapshot slots: at most 3 are in flight
        nslot = 0
        deferred = None                # (point, snapshot) of a point whose LAST batch is still running
        stopped = False

        def close(point, entry):
            """take a finished point's counters; True if the sweep ends here (links.py:339-341)"""
            entry[2].synchronize()
            c = entry[1].numpy()
            grand.add_(entry[0])
            BERs[point] = c[0] / c[2]
            return bool(stop_early and c[0] < err_min)

        for i, snr in enumerate(SNRs):
            tot = torch.zeros(3, dtype=torch.int64, device="cuda")          # bit errors, frame errors, bits sent
            hist = []                                                        # (device snapshot, pinned copy, event)
            bits_known = 0
            while True:
                msg, y, nv = self.make_batch(float(snr), batch_index, torch)
                batch_index += 1
                local = torch.zeros(3, dtype=torch.int64, device="cuda")
                self.receive_decode_count(msg, y, nv, local, torch)
                local[2] = msg.numel()
                parallel.allreduce_counters(local)
                tot = tot + local
                host = pinned[nslot]
                nslot = (nslot + 1) % pinned.shape[0]
                host.copy_(tot, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                hist.append((tot, host, ev))
                bits_known += bits_per_batch
                if deferred is not None:
                    # the previous point's last batch ran while this point's first batch was being issued: the GPU never
                    # waits for the host between points (this batch is discarded if the sweep ends at that point)
                    point, entry = deferred
                    deferred = None
                    if close(point, entry):
                        stopped = True
                        break
                if len(hist) >= 2:                                           # batch b-1, while batch b runs
                    hist[-2][2].synchronize()
                    if int(hist[-2][1][0]) >= err_min:
                        if close(i, hist[-2]):
                            stopped = True
                        break
                if bits_known >= send_max:
                    if overlap_points:
                        deferred = (i, hist[-1])
                    elif close(i, hist[-1]):
                        stopped = True
                    break
            if stopped:
                break
        if deferred is not None and not stopped:
            close(*deferred)
        if return_counters:
            return BERs, grand
        return BERs


def turbo_link_tx(trellis, interleaver, frames, frame_bits, seed, first_frame, noise_sigma):
    """Device-side TX chain of a turbo-coded BPSK link (cpb_turbo_link_tx): random message -> turbo_encode(msg, trellis,
    trellis, interleaver) -> 2x-1 -> + noise_sigma * N(0,1), for `frames` frames starting at GLOBAL frame `first_frame`.

    Returns (msg uint8, sys, par1, par2 float32), each (frames, frame_bits), CUDA tensors: what map_decode / turbo_decode
    take.  The streams depend only on (seed, global frame index)."""
    import ctypes as C
    from .channelcoding.convcode import _trellis_handle
    from .channelcoding.turbo import _checked_perm
    torch = _lib.require_cuda()
    N = int(frame_bits)
    perm = torch.from_numpy(_checked_perm(interleaver, N)).cuda()
    msg = torch.empty((int(frames), N), dtype=torch.uint8, device="cuda")
    ys, y1, y2 = (torch.empty((int(frames), N), dtype=torch.float32, device="cuda") for _ in range(3))
    rc = _lib.load().cpb_turbo_link_tx(_trellis_handle(trellis), _lib.ptr(perm), C.c_int64(int(frames)), C.c_int64(N),
                                       C.c_uint64(int(seed) & ((1 << 64) - 1)), C.c_int64(int(first_frame)),
                                       C.c_float(float(noise_sigma)), _lib.ptr(msg), _lib.ptr(ys), _lib.ptr(y1), _lib.ptr(y2),
                                       _lib.stream_ptr(torch))
    _lib.check(rc, "turbo_link_tx")
    return msg, ys, y1, y2


class TurboLinkGPU:
    """Batched rate-1/3 turbo link over BPSK-AWGN on the GPU(s): frames generated (cpb_turbo_link_tx), decoded
    (cpb_turbo_decode) and counted (cpb_count_errors) on the device; the error counters are the only thing all-reduced."""

    def __init__(self, trellis, interleaver, frame_bits, frames_per_batch=1024, iterations=6, seed=0):
        self.trellis, self.interleaver = trellis, interleaver
        self.frame_bits, self.frames, self.iterations, self.seed = int(frame_bits), int(frames_per_batch), int(iterations), int(seed)

    def noise_variance(self, ebn0_db):
        return 1.0 / (2.0 * (1.0 / 3.0) * 10 ** (ebn0_db / 10.0))

    def make_batch(self, ebn0_db, batch_index):
        rank, world, _ = parallel.world()
        first = parallel.batch_first_frame(batch_index, self.frames, rank, max(world, 1))
        s2 = self.noise_variance(ebn0_db)
        return turbo_link_tx(self.trellis, self.interleaver, self.frames, self.frame_bits, self.seed, first, math.sqrt(s2)) + (s2,)

    def decode_count(self, msg, ys, y1, y2, s2, counters, torch):
        import ctypes as C
        from .channelcoding import turbo_decode_batch
        dec = turbo_decode_batch(ys, y1, y2, self.trellis, s2, self.iterations, self.interleaver)
        L = msg.shape[1]
        rc = _lib.load().cpb_count_errors(_lib.ptr(dec), _lib.ptr(msg), C.c_int64(msg.shape[0]), C.c_int64(L), C.c_int64(L),
                                          C.c_int64(L), _lib.ptr(counters), _lib.stream_ptr(torch))
        _lib.check(rc, "count_errors")
        return dec

    def link_performance(self, EbN0s, send_max, err_min):
        """BER per Eb/N0 (dB): a point ends when the global counters reach `err_min` bit errors or `send_max` bits."""
        torch = _lib.require_cuda()
        BERs = np.zeros(len(EbN0s))
        batch_index = 0
        for i, e in enumerate(EbN0s):
            tot = torch.zeros(3, dtype=torch.int64, device="cuda")
            while True:
                msg, ys, y1, y2, s2 = self.make_batch(float(e), batch_index)
                batch_index += 1
                local = torch.zeros(3, dtype=torch.int64, device="cuda")
                self.decode_count(msg, ys, y1, y2, s2, local, torch)
                local[2] = msg.numel()
                parallel.allreduce_counters(local)
                tot += local
                c = tot.cpu().numpy()
                if not parallel.stop_rule(c, send_max, err_min):
                    break
            BERs[i] = c[0] / c[2]
        return BERs


def idd_decoder(detector, decoder, decision, n_it):
    """Iterative detection and decoding for a coded MIMO link (links.py:345-407): returns the 6-argument decoder LinkModel
    calls.  Each iteration decodes the current a-priori LLRs, hands the extrinsic part to `detector(y_i, H_i, constellation,
    noise_var, a_priori_i)` vector by vector, and keeps the detector's extrinsic for the next round; `decision` maps the
    final LLRs to bits."""
    def decode(y, h, constellation, noise_var, a_priori, bits_per_send):
        to_decoder = np.array(a_priori, dtype=float)
        nb_vect = h.shape[0]
        to_detector = np.zeros_like(to_decoder)
        for _ in range(n_it):
            to_detector = decoder(to_decoder) - to_decoder
            for i in range(nb_vect):
                sl = slice(i * bits_per_send, (i + 1) * bits_per_send)
                to_decoder[sl] = detector(y[i], h[i], constellation, noise_var, to_detector[sl])
            to_decoder -= to_detector
        return decision(to_decoder + to_detector)
    return decode
