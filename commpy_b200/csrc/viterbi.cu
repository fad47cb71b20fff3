// Viterbi decoding on sm_100a -- replaces the loop nest of commpy/channelcoding/convcode.py:661-749
// (viterbi_decode), :590-657 (_acs_traceback), :575-587 (_compute_branch_metrics), :561-572 (_where_c).
//
// Exact semantics kept from the reference:
//   * T = int((L+M)/k) - 1 trellis steps, zero / -1 padding past the received data;
//   * add-compare-select picks the FIRST minimum in (prev_state asc, input asc) order;
//   * the symbol of step q is read on the survivor path that starts at the lowest-index best state of
//     step min(q + D - 2, T) (D = tb_depth): one sliding-window traceback per output symbol.
//
// Two kernel families:
//   viterbi_fast_kernel     k=1, n=2 feed-forward shift-register codes with 64 states (K=7): one thread owns
//                           all 64 path metrics of one frame (int32 fixed-point LLR metrics) or of TWO frames
//                           (hard decision, u16x2-packed).  Metrics are kept in "key form"
//                           metric*64 + state_index so that a single VIADDMNMX does add+compare+select with the
//                           reference tie rule, the survivor bit is the key's LSB and the best state of a step is
//                           a VIMNMX3 tree.  Survivors live in a shared-memory ring of D+9 steps per frame.
//   viterbi_generic_kernel  any trellis (k<=4, n<=4, S<=256): table driven, one thread per frame, fp32 metrics
//                           in shared memory, survivors in a global scratch buffer.
// Both use the same block traceback: one long traceback per 16/32 steps plus a per-window fallback whenever
// the long path does not pass through that window's own best state, which reproduces the reference's
// per-step traceback bit for bit.
#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace cpb;

struct cpbTrellis {
    int k, n, M, S, I;
    std::vector<int32_t> next_state, output;   // host copies, S x I
    int32_t *pred_dev = nullptr;               // S*I entries: prev_state | input<<8 | output<<16, (p asc, u asc)
    int32_t *next_dev = nullptr, *out_dev = nullptr;   // S x I tables on the device (BCJR)
    int fast_id = 0;
    int device = 0;
};

// ------------------------------------------------------------------------------------------------
// compile-time description of a rate-1/2 feed-forward code in CommPy's Trellis convention
// (convcode.py:195-255, polynomial_format='MSB'): generator bit b multiplies delay b (bit 0 = current
// input); state bit (M-b) holds delay b; output symbol = (parity(G0)<<1) | parity(G1).
// ------------------------------------------------------------------------------------------------
template <int M_, uint32_t G0_, uint32_t G1_>
struct FFCode {
    static constexpr int M = M_;
    static constexpr int S = 1 << M_;
    static constexpr uint32_t G0 = G0_, G1 = G1_;
    __host__ __device__ static constexpr int parity(uint32_t v)
    {
        v ^= v >> 16; v ^= v >> 8; v ^= v >> 4; v ^= v >> 2; v ^= v >> 1;
        return (int)(v & 1u);
    }
    __host__ __device__ static constexpr uint32_t regs(int s, int u)
    {
        uint32_t v = (uint32_t)u;
        for (int b = 1; b <= M_; ++b) v |= (uint32_t)((s >> (M_ - b)) & 1) << b;
        return v;
    }
    __host__ __device__ static constexpr int out(int s, int u)
    {
        return (parity(regs(s, u) & G0_) << 1) | parity(regs(s, u) & G1_);
    }
};

using Code133_171 = FFCode<6, 0133, 0171>;   // the standard K=7 code (octal 133,171)
using Code171_133 = FFCode<6, 0171, 0133>;
using Code5_43 = FFCode<6, 5, 43>;           // what Trellis makes of DECIMAL (133,171): wifi80211.py:49 quirk

template <class CODE>
static bool code_matches(const cpbTrellis &t)
{
    if (t.k != 1 || t.n != 2 || t.M != CODE::M || t.S != CODE::S) return false;
    for (int s = 0; s < CODE::S; ++s)
        for (int u = 0; u < 2; ++u) {
            if (t.next_state[s * 2 + u] != ((u << (CODE::M - 1)) | (s >> 1))) return false;
            if (t.output[s * 2 + u] != CODE::out(s, u)) return false;
        }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Fast path
// ------------------------------------------------------------------------------------------------
namespace fast {

constexpr int TBB = 16;            // windows per traceback block
constexpr int QBITS = 19;          // |quantised LLR| <= 2^19
constexpr int QMAX = 1 << QBITS;

struct Params {
    const void *coded;
    int64_t n_in;
    int64_t batch;
    int L, T, D, R;
    int mode;                // CPB_VITERBI_*
    const uint32_t *amax_bits;   // float input: bits of max |x| over the call (device)
    uint8_t *out;
    int out_vec16;           // 1: rows of out are 16-byte aligned
};

template <int PACK> struct KeyOps;
template <> struct KeyOps<2> {   // two frames per register, u16 halves
    static constexpr uint32_t IDX_MASK = 0x003F003Fu;
    static constexpr uint32_t LSB = 0x00010001u;
    __device__ static __forceinline__ uint32_t addmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_u16x2(a, b, c); }
    __device__ static __forceinline__ uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
    __device__ static __forceinline__ uint32_t min2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
    __device__ static __forceinline__ uint32_t idx(int s) { return (uint32_t)s | ((uint32_t)s << 16); }
};
template <> struct KeyOps<1> {
    static constexpr uint32_t IDX_MASK = 0x3Fu;
    static constexpr uint32_t LSB = 1u;
    __device__ static __forceinline__ uint32_t addmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_u32(a, b, c); }
    __device__ static __forceinline__ uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u32(a, b, c); }
    __device__ static __forceinline__ uint32_t min2(uint32_t a, uint32_t b) { return min(a, b); }
    __device__ static __forceinline__ uint32_t idx(int s) { return (uint32_t)s; }
};

// one trellis step on register-resident keys: Kn <- ACS(K, Bm); W <- survivor bits; returns min key(s)
template <class CODE, int PACK>
__device__ __forceinline__ uint32_t acs_step(const uint32_t (&K)[64], uint32_t (&Kn)[64], const uint32_t (&Bm)[4],
                                             uint32_t (&W)[2 * PACK])
{
    using OPS = KeyOps<PACK>;
    constexpr int H = CODE::S / 2;
    constexpr int WSH = (PACK == 2) ? 4 : 5;        // states per survivor word: 16 (packed) or 32
#pragma unroll
    for (int w = 0; w < 2 * PACK; ++w) W[w] = 0;
#pragma unroll
    for (int l = 0; l < H; ++l) {
        // predecessors 2l (survivor bit 0) and 2l+1 (bit 1); new state l <- input 0, l+H <- input 1
        const uint32_t c10 = K[2 * l + 1] + Bm[CODE::out(2 * l + 1, 0)];
        const uint32_t m0 = OPS::addmin(K[2 * l], Bm[CODE::out(2 * l, 0)], c10);
        const uint32_t c11 = K[2 * l + 1] + Bm[CODE::out(2 * l + 1, 1)];
        const uint32_t m1 = OPS::addmin(K[2 * l], Bm[CODE::out(2 * l, 1)], c11);
        Kn[l] = (m0 & ~OPS::IDX_MASK) | OPS::idx(l);
        Kn[l + H] = (m1 & ~OPS::IDX_MASK) | OPS::idx(l + H);
        W[l >> WSH] += (m0 & OPS::LSB) << (l & ((1 << WSH) - 1));
        W[(l + H) >> WSH] += (m1 & OPS::LSB) << (l & ((1 << WSH) - 1));
    }
    // best state(s): lowest key = lowest metric, ties -> lowest state index (np.argmin, convcode.py:645)
    uint32_t r[22];
#pragma unroll
    for (int i = 0; i < 21; ++i) r[i] = OPS::min3(Kn[3 * i], Kn[3 * i + 1], Kn[3 * i + 2]);
    r[21] = Kn[63];
    uint32_t q[8];
#pragma unroll
    for (int i = 0; i < 7; ++i) q[i] = OPS::min3(r[3 * i], r[3 * i + 1], r[3 * i + 2]);
    q[7] = r[21];
    const uint32_t a = OPS::min3(q[0], q[1], q[2]);
    const uint32_t b = OPS::min3(q[3], q[4], q[5]);
    return OPS::min3(a, b, OPS::min2(q[6], q[7]));
}

// shared-memory survivor ring, thread-private columns: words [slot][w][tid], best [slot][tid]
template <int PACK>
struct Ring {
    uint32_t *w;
    uint16_t *best;
    int bd, tid, R;
    __device__ __forceinline__ void store(int slot, const uint32_t (&W)[2 * PACK], uint32_t bestv)
    {
#pragma unroll
        for (int i = 0; i < 2 * PACK; ++i) w[(slot * 2 * PACK + i) * bd + tid] = W[i];
        best[slot * bd + tid] = (uint16_t)bestv;
    }
    __device__ __forceinline__ int get_best(int slot, int fi) const { return (best[slot * bd + tid] >> (8 * fi)) & 63; }
    __device__ __forceinline__ int get_dec(int slot, int s, int fi) const
    {
        if (PACK == 2) return (w[(slot * 4 + (s >> 4)) * bd + tid] >> ((s & 15) + 16 * fi)) & 1;
        return (w[(slot * 2 + (s >> 5)) * bd + tid] >> (s & 31)) & 1;
    }
    __device__ __forceinline__ int dec_slot(int slot) const { return slot == 0 ? R - 1 : slot - 1; }
};

// Traceback for the windows t' in (ts, te].  slot_te = ring slot of step te.
// Bit p (0-based) of the frame is the input of step q = p+1; on a path whose state at step tau is s,
// that input is bit (tau - q) of ... the state holds the last M inputs, newest in the MSB, so the input of
// step tau-(M-1) is the LSB of s (App. A.1-9).
template <class CODE, int PACK>
__device__ __forceinline__ void tb_block(const Ring<PACK> &ring, int ts, int te, bool final_blk, int slot_te, int D,
                                         int L, uint8_t *const (&outp)[PACK], const bool (&valid)[PACK], int out_vec16)
{
    constexpr int M = CODE::M, S = CODE::S;
    const int p0 = ts - D + 2;                 // first bit of this block
    const int tau_min = ts - D + 3 + (M - 1);
    const int q_hi = final_blk ? te : te - D + 2;
#pragma unroll
    for (int fi = 0; fi < PACK; ++fi) {
        unsigned long long acc = 0ull;
        uint32_t cons = 0;
        int s = ring.get_best(slot_te, fi);
        int slot = slot_te;
        for (int tau = te; tau >= tau_min; --tau) {
            if (tau > ts) cons |= (uint32_t)(s == ring.get_best(slot, fi)) << (tau - ts - 1);
            const int q = tau - (M - 1);
            if (q >= 1 && q <= q_hi) acc |= (unsigned long long)(s & 1) << (q - 1 - p0);
            s = ((s << 1) & (S - 1)) | ring.get_dec(slot, s, fi);
            slot = ring.dec_slot(slot);
        }
        // windows whose own best state is not on the long path: individual (D-1)-step traceback
        const int nwin = te - ts;
        uint32_t todo = ~cons & ((nwin >= 32) ? 0xffffffffu : ((1u << nwin) - 1u));
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const int tp = ts + 1 + j;
            int sl = slot_te - (te - tp);
            if (sl < 0) sl += ring.R;
            int s2 = ring.get_best(sl, fi);
            for (int i = 0; i < D - 2 - (M - 1); ++i) {
                s2 = ((s2 << 1) & (S - 1)) | ring.get_dec(sl, s2, fi);
                sl = ring.dec_slot(sl);
            }
            acc = (acc & ~(1ull << j)) | ((unsigned long long)(s2 & 1) << j);
        }
        if (!valid[fi]) continue;
        uint8_t *o = outp[fi] + p0;
        if (!final_blk && out_vec16) {
            const uint32_t b16 = (uint32_t)acc & 0xffffu;
            uint4 v;
            v.x = (((b16 >> 0) & 15u) * 0x00204081u) & 0x01010101u;
            v.y = (((b16 >> 4) & 15u) * 0x00204081u) & 0x01010101u;
            v.z = (((b16 >> 8) & 15u) * 0x00204081u) & 0x01010101u;
            v.w = (((b16 >> 12) & 15u) * 0x00204081u) & 0x01010101u;
            *reinterpret_cast<uint4 *>(o) = v;
        } else {
            const int cnt = final_blk ? (L - p0) : TBB;
            for (int i = 0; i < cnt; ++i) o[i] = (uint8_t)((acc >> i) & 1ull);
        }
    }
}

template <class CODE, int PACK>
__global__ void __launch_bounds__(128) viterbi_fast_kernel(const Params p)
{
    using OPS = KeyOps<PACK>;
    constexpr int S = CODE::S, M = CODE::M;
    static_assert(S == 64, "fast path is written for 64 states");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int bd = blockDim.x, tid = threadIdx.x;
    Ring<PACK> ring;
    ring.bd = bd; ring.tid = tid; ring.R = p.R;
    ring.w = reinterpret_cast<uint32_t *>(smem_raw);
    ring.best = reinterpret_cast<uint16_t *>(smem_raw + (size_t)p.R * 2 * PACK * bd * sizeof(uint32_t));
    uint4 *lut = reinterpret_cast<uint4 *>(smem_raw + (size_t)p.R * 2 * PACK * bd * sizeof(uint32_t) +
                                           (((size_t)p.R * bd * sizeof(uint16_t) + 15) & ~(size_t)15));
    if (PACK == 2) {
        // hard-decision branch metrics of two frames: entry idx = r0A | r1A<<1 | r0B<<2 | r1B<<3,
        // component o = Hamming distance to output symbol o (convcode.py:579), times 64, frame B in the high half
        if (tid < 16) {
            const int a = ((tid & 1) << 1) | ((tid >> 1) & 1), b = (((tid >> 2) & 1) << 1) | ((tid >> 3) & 1);
            uint32_t e[4];
            for (int o = 0; o < 4; ++o) e[o] = ((uint32_t)__popc(o ^ a) << 6) | ((uint32_t)__popc(o ^ b) << 22);
            lut[tid] = make_uint4(e[0], e[1], e[2], e[3]);
        }
        __syncthreads();
    }

    // frames of this thread
    const int64_t base = (int64_t)blockIdx.x * bd * PACK;
    int64_t fr[PACK];
    bool valid[PACK];
    uint8_t *outp[PACK];
#pragma unroll
    for (int fi = 0; fi < PACK; ++fi) {
        fr[fi] = base + (int64_t)fi * bd + tid;
        valid[fi] = fr[fi] < p.batch;
        if (!valid[fi]) fr[fi] = p.batch - 1;
        outp[fi] = p.out + fr[fi] * (int64_t)p.L;
    }

    // float input: power-of-two scale so that the largest |value| of the call maps to ~2^19
    float scale = 1.0f, padq = 0.0f;
    if (PACK == 1) {
        float amax = __uint_as_float(*p.amax_bits);
        if (p.mode == CPB_VITERBI_SOFT) amax = fminf(amax, 500.0f);
        amax = fminf(fmaxf(amax, 1e-30f), 3.0e38f);
        scale = exp2f(floorf(log2f((float)QMAX / amax)));
        scale = fminf(scale, 1.0e30f);
        padq = (p.mode == CPB_VITERBI_UNQUANTIZED) ? -1.0f : 0.0f;   // convcode.py:727-732
    }

    uint32_t K[64], Kn[64];
    {
        // pm[0] = 0, every other state "infinite" (convcode.py:705-706): a finite sentinel larger than any
        // metric a path starting in state 0 can lose against (n*M*max branch metric) behaves identically.
        const uint32_t big = (PACK == 2) ? (16u << 6) : ((uint32_t)(2 * M * QMAX + 1) << 6);
#pragma unroll
        for (int s = 0; s < 64; ++s) {
            const uint32_t v = (s == 0) ? 0u : big;
            K[s] = (PACK == 2) ? ((v | (v << 16)) | OPS::idx(s)) : (v | OPS::idx(s));
        }
    }

    const unsigned char *c8 = reinterpret_cast<const unsigned char *>(p.coded);
    const float *cf = reinterpret_cast<const float *>(p.coded);

    // received pair of step tau -> the four branch metrics Bm[o] (times 64)
    auto load_raw = [&](int tau, uint32_t &ra, uint32_t &rb) {
        // hard: ra = idx bits for frame A/B ; float: ra, rb = bits of r0, r1
        if (PACK == 2) {
            uint32_t idx = 0;
            if (tau <= p.L) {
#pragma unroll
                for (int fi = 0; fi < PACK; ++fi) {
                    const unsigned char *q = c8 + fr[fi] * p.n_in + 2 * (int64_t)(tau - 1);
                    idx |= ((uint32_t)(__ldg(q) & 1u) | ((uint32_t)(__ldg(q + 1) & 1u) << 1)) << (2 * fi);
                }
            }
            ra = idx; rb = 0;
        } else {
            float r0 = padq, r1 = padq;
            if (tau <= p.L) {
                const float *q = cf + fr[0] * p.n_in + 2 * (int64_t)(tau - 1);
                r0 = __ldg(q); r1 = __ldg(q + 1);
            }
            ra = __float_as_uint(r0); rb = __float_as_uint(r1);
        }
    };
    auto make_bm = [&](uint32_t ra, uint32_t rb, uint32_t (&Bm)[4]) {
        if (PACK == 2) {
            const uint4 e = lut[ra];
            Bm[0] = e.x; Bm[1] = e.y; Bm[2] = e.z; Bm[3] = e.w;
        } else {
            float r0 = __uint_as_float(ra), r1 = __uint_as_float(rb);
            if (p.mode == CPB_VITERBI_SOFT) {          // convcode.py:718-719
                r0 = fminf(fmaxf(r0, -500.0f), 500.0f);
                r1 = fminf(fmaxf(r1, -500.0f), 500.0f);
            }
            int q0 = __float2int_rn(fminf(fmaxf(r0 * scale, -(float)QMAX), (float)QMAX));
            int q1 = __float2int_rn(fminf(fmaxf(r1 * scale, -(float)QMAX), (float)QMAX));
            // -log-likelihood of code bit c given value r, up to a per-step constant (convcode.py:581-587):
            // c = 0 costs max(q,0), c = 1 costs max(-q,0)
            const uint32_t z0 = (uint32_t)max(q0, 0) << 6, o0 = (uint32_t)max(-q0, 0) << 6;
            const uint32_t z1 = (uint32_t)max(q1, 0) << 6, o1 = (uint32_t)max(-q1, 0) << 6;
            Bm[0] = z0 + z1; Bm[1] = z0 + o1; Bm[2] = o0 + z1; Bm[3] = o0 + o1;
        }
    };

    int slot = 0;
    int next_te = p.D - 2 + TBB;
    uint32_t W[2 * PACK];

    auto finish_step = [&](int tau, uint32_t mn, uint32_t (&Kc)[64]) {
        const uint32_t bestv = (PACK == 2) ? ((mn & 63u) | (((mn >> 16) & 63u) << 8)) : (mn & 63u);
        ring.store(slot, W, bestv);
        if ((tau & 15) == 0) {          // renormalise: subtract the minimum metric from every key
            const uint32_t sub = mn & ~OPS::IDX_MASK;
#pragma unroll
            for (int s = 0; s < 64; ++s) Kc[s] -= sub;
        }
        if (tau == p.T) {
            tb_block<CODE, PACK>(ring, next_te - TBB, tau, true, slot, p.D, p.L, outp, valid, p.out_vec16);
        } else if (tau == next_te) {
            tb_block<CODE, PACK>(ring, next_te - TBB, tau, false, slot, p.D, p.L, outp, valid, p.out_vec16);
            next_te += TBB;
        }
        slot = (slot + 1 == p.R) ? 0 : slot + 1;
    };

    uint32_t ra0, rb0, ra1, rb1;
    load_raw(1, ra0, rb0);
    load_raw(2, ra1, rb1);
    for (int tau = 1; tau <= p.T; tau += 2) {
        uint32_t na0, nb0, na1, nb1;
        load_raw(tau + 2, na0, nb0);      // software prefetch of the next pair of steps
        load_raw(tau + 3, na1, nb1);
        uint32_t Bm[4];
        make_bm(ra0, rb0, Bm);
        uint32_t mn = acs_step<CODE, PACK>(K, Kn, Bm, W);
        finish_step(tau, mn, Kn);
        if (tau + 1 <= p.T) {
            make_bm(ra1, rb1, Bm);
            mn = acs_step<CODE, PACK>(Kn, K, Bm, W);
            finish_step(tau + 1, mn, K);
        }
        ra0 = na0; rb0 = nb0; ra1 = na1; rb1 = nb1;
    }
}

// max |x| over a float buffer, as uint bits (non-negative floats order like unsigned ints)
__global__ void absmax_kernel(const float *__restrict__ x, int64_t n, int clip500, uint32_t *out_bits)
{
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = fabsf(x[i]);
        if (!(v <= 3.0e38f)) v = 3.0e38f;      // inf / nan
        if (clip500) v = fminf(v, 500.0f);
        m = fmaxf(m, v);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

static size_t smem_bytes(int R, int bd, int pack)
{
    size_t ring_w = (size_t)R * 2 * pack * bd * sizeof(uint32_t);
    size_t ring_b = (((size_t)R * bd * sizeof(uint16_t)) + 15) & ~(size_t)15;
    return ring_w + ring_b + 16 * sizeof(uint4);
}

template <class CODE, int PACK>
static int launch(const Params &p, int bd, cudaStream_t st)
{
    const size_t smem = smem_bytes(p.R, bd, PACK);
    auto kern = viterbi_fast_kernel<CODE, PACK>;
    CPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = ceil_div(p.batch, (int64_t)bd * PACK);
    kern<<<(unsigned)grid, bd, smem, st>>>(p);
    CPB_LAUNCH_CHECK();
    return CPB_OK;
}

}  // namespace fast

// ------------------------------------------------------------------------------------------------
// Generic table-driven path
// ------------------------------------------------------------------------------------------------
namespace gen {

constexpr int BD = 64;
constexpr int TBB = 32;

struct Params {
    const void *coded;
    int in_dtype;
    int64_t n_in;
    int64_t frame0;       // first frame of this chunk
    int nframes;          // frames in this chunk
    int64_t stride;       // frame stride of the scratch planes (>= nframes)
    const int32_t *pred;
    int k, n, S, I;
    int L, T, D, mode;
    uint8_t *winners;     // [(T+1)*S][stride]
    uint8_t *best;        // [T+1][stride]
    uint8_t *out;
};

__global__ void __launch_bounds__(BD) viterbi_generic_kernel(const Params p)
{
    extern __shared__ float sm[];
    const int tid = threadIdx.x;
    const int S = p.S, I = p.I, n = p.n, k = p.k;
    const int NB = 1 << n;
    float *pm0 = sm;                       // [S][BD]
    float *pm1 = sm + (size_t)S * BD;      // [S][BD]
    float *bmv = sm + (size_t)2 * S * BD;  // [NB][BD]
    const int f = blockIdx.x * BD + tid;
    const bool valid = f < p.nframes;
    const int64_t frame = p.frame0 + (valid ? f : 0);
    const unsigned char *c8 = reinterpret_cast<const unsigned char *>(p.coded) + frame * p.n_in;
    const float *cf = reinterpret_cast<const float *>(p.coded) + frame * p.n_in;
    const int64_t fcol = valid ? f : 0;

    for (int s = 0; s < S; ++s) pm0[s * BD + tid] = (s == 0) ? 0.0f : INFINITY;     // convcode.py:705-706
    float *po = pm0, *pn = pm1;
    const int64_t Lk = p.L / k;
    const float padv = (p.mode == CPB_VITERBI_UNQUANTIZED) ? -1.0f : 0.0f;

    for (int tau = 1; tau <= p.T; ++tau) {
        float r[4];
        for (int j = 0; j < n; ++j) {
            float v = padv;
            if (tau <= Lk) {
                const int64_t e = (int64_t)(tau - 1) * n + j;
                v = (p.in_dtype == CPB_U8) ? (float)(c8[e] & 1u) : cf[e];
            }
            if (p.mode == CPB_VITERBI_SOFT) v = fminf(fmaxf(v, -500.0f), 500.0f);
            r[j] = v;
        }
        for (int c = 0; c < NB; ++c) {
            float acc = 0.0f;
            for (int j = 0; j < n; ++j) {
                const int cj = (c >> (n - 1 - j)) & 1;           // MSB first
                if (p.mode == CPB_VITERBI_HARD) acc += (((int)r[j]) ^ cj) ? 1.0f : 0.0f;
                else if (p.mode == CPB_VITERBI_SOFT) acc += cj ? fmaxf(-r[j], 0.0f) : fmaxf(r[j], 0.0f);
                else { const float d = r[j] - (float)(2 * cj - 1); acc += d * d; }
            }
            bmv[c * BD + tid] = acc;
        }
        float mn = INFINITY; int arg = 0;
        for (int s = 0; s < S; ++s) {
            float bestm = 0.0f; int bi = 0;
            for (int i = 0; i < I; ++i) {
                const int e = __ldg(&p.pred[s * I + i]);
                const float m = po[(e & 0xff) * BD + tid] + bmv[((e >> 16) & 0xff) * BD + tid];
                if (i == 0 || m < bestm) { bestm = m; bi = i; }
            }
            pn[s * BD + tid] = bestm;
            if (valid) p.winners[((int64_t)tau * S + s) * p.stride + fcol] = (uint8_t)bi;
            if (s == 0 || bestm < mn) { mn = bestm; arg = s; }
        }
        if (valid) p.best[(int64_t)tau * p.stride + fcol] = (uint8_t)arg;
        if (mn < INFINITY && mn != 0.0f)
            for (int s = 0; s < S; ++s) pn[s * BD + tid] -= mn;     // exact for the integer metrics of 'hard'
        float *t_ = po; po = pn; pn = t_;
    }
    if (!valid) return;

    // traceback: symbol of step q comes from the path started at best[min(q + D - 2, T)]
    uint8_t *o = p.out + frame * (int64_t)p.L;
    auto emit = [&](int q, int u) {
        for (int b = 0; b < k; ++b) {
            const int64_t pos = (int64_t)(q - 1) * k + b;
            if (pos < p.L) o[pos] = (uint8_t)((u >> (k - 1 - b)) & 1);
        }
    };
    auto prev = [&](int tau, int s, int &u) {
        const int i = p.winners[((int64_t)tau * S + s) * p.stride + fcol];
        const int e = __ldg(&p.pred[s * I + i]);
        u = (e >> 8) & 0xff;
        return e & 0xff;
    };
    auto bestat = [&](int tau) { return (int)p.best[(int64_t)tau * p.stride + fcol]; };

    int ts = p.D - 2;
    while (ts < p.T) {
        int te = ts + TBB;
        bool fin = false;
        if (te >= p.T) { te = p.T; fin = true; }
        int s = bestat(te);
        uint32_t cons = 0;
        const int tau_min = max(1, ts - p.D + 3);
        for (int tau = te; tau >= tau_min; --tau) {
            if (tau > ts && s == bestat(tau)) cons |= 1u << (tau - ts - 1);
            int u;
            const int pr = prev(tau, s, u);
            if (fin || tau <= te - p.D + 2) emit(tau, u);
            s = pr;
        }
        for (int tp = ts + 1; tp < te; ++tp) {
            if ((cons >> (tp - ts - 1)) & 1u) continue;
            int s2 = bestat(tp), u = 0;
            for (int tau = tp; tau >= tp - p.D + 2; --tau) {
                if (tau < 1) break;
                s2 = prev(tau, s2, u);
            }
            if (tp - p.D + 2 >= 1) emit(tp - p.D + 2, u);
        }
        ts = te;
    }
}

}  // namespace gen

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int cpb_trellis_create(const int32_t *next_state, const int32_t *output, int k, int n, int total_memory,
                       int number_states, cpbTrellis **out)
{
    if (!next_state || !output || !out || k < 1 || k > 4 || n < 1 || n > 4 || number_states < 1 ||
        number_states > 256 || total_memory < 0)
        return CPB_EINVAL;
    cpbTrellis *t = new cpbTrellis();
    t->k = k; t->n = n; t->M = total_memory; t->S = number_states; t->I = 1 << k;
    const int S = t->S, I = t->I;
    t->next_state.assign(next_state, next_state + S * I);
    t->output.assign(output, output + S * I);
    std::vector<int32_t> pred(S * I, 0), cnt(S, 0);
    for (int p = 0; p < S; ++p)
        for (int u = 0; u < I; ++u) {
            const int s = next_state[p * I + u], o = output[p * I + u];
            if (s < 0 || s >= S || o < 0 || o >= (1 << n) || cnt[s] >= I) { delete t; return CPB_ETRELLIS; }
            pred[s * I + cnt[s]++] = p | (u << 8) | (o << 16);     // (p asc, u asc): convcode.py:561-572
        }
    for (int s = 0; s < S; ++s)
        if (cnt[s] != I) { delete t; return CPB_ETRELLIS; }
    cudaGetDevice(&t->device);
    const size_t bytes = sizeof(int32_t) * S * I;
    if (cudaMalloc(&t->pred_dev, bytes) != cudaSuccess || cudaMalloc(&t->next_dev, bytes) != cudaSuccess ||
        cudaMalloc(&t->out_dev, bytes) != cudaSuccess) {
        record_cuda_error(cudaGetLastError(), "cudaMalloc(trellis tables)", __FILE__, __LINE__);
        cpb_trellis_destroy(t);
        return CPB_ECUDA;
    }
    cudaMemcpy(t->pred_dev, pred.data(), bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(t->next_dev, t->next_state.data(), bytes, cudaMemcpyHostToDevice);
    cudaError_t e = cudaMemcpy(t->out_dev, t->output.data(), bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        record_cuda_error(e, "cudaMemcpy(trellis tables)", __FILE__, __LINE__);
        cpb_trellis_destroy(t);
        return CPB_ECUDA;
    }
    if (code_matches<Code133_171>(*t)) t->fast_id = 1;
    else if (code_matches<Code171_133>(*t)) t->fast_id = 2;
    else if (code_matches<Code5_43>(*t)) t->fast_id = 3;
    *out = t;
    return CPB_OK;
}

int cpb_trellis_destroy(cpbTrellis *t)
{
    if (!t) return CPB_OK;
    if (t->pred_dev) cudaFree(t->pred_dev);
    if (t->next_dev) cudaFree(t->next_dev);
    if (t->out_dev) cudaFree(t->out_dev);
    delete t;
    return CPB_OK;
}

int cpb_trellis_fast_path(const cpbTrellis *t) { return t ? t->fast_id : 0; }

}  // extern "C"

// tables for the BCJR kernels (bcjr.cu)
const int32_t *cpb_trellis_next_dev(const cpbTrellis *t) { return t->next_dev; }
const int32_t *cpb_trellis_out_dev(const cpbTrellis *t) { return t->out_dev; }
const int32_t *cpb_trellis_pred_dev(const cpbTrellis *t) { return t->pred_dev; }
void cpb_trellis_dims(const cpbTrellis *t, int *k, int *n, int *S) { *k = t->k; *n = t->n; *S = t->S; }

static int resolve_depth(const cpbTrellis *t, int64_t L, int tb_depth)
{
    if (tb_depth <= 0) {
        int64_t d = 5 * (int64_t)t->M;          // convcode.py:701-702
        if (d > L) d = L;
        return (int)d;
    }
    return tb_depth;
}

static bool use_fast(const cpbTrellis *t, int D, int mode, int in_dtype)
{
    if (t->fast_id == 0) return false;
    if (D < t->M + 1 || D > 48) return false;
    if (mode == CPB_VITERBI_HARD) return in_dtype == CPB_U8;
    return in_dtype == CPB_F32;
}

static size_t generic_chunk(const cpbTrellis *t, int64_t batch, int64_t T, int64_t *stride)
{
    // survivors: (T+1)*(S+1) bytes per frame; keep one chunk under ~1.5 GB
    const double per_frame = (double)(T + 1) * (t->S + 1);
    int64_t chunk = (int64_t)(1.5e9 / per_frame);
    chunk = (chunk / gen::BD) * gen::BD;
    if (chunk < gen::BD) chunk = gen::BD;
    const int64_t need = ceil_div(batch, gen::BD) * gen::BD;
    if (chunk > need) chunk = need;
    *stride = chunk;
    return (size_t)((T + 1) * (int64_t)(t->S + 1) * chunk);
}

extern "C" {

int cpb_viterbi_sizes(const cpbTrellis *t, int64_t n_in, int64_t *L, int64_t *T)
{
    if (!t || n_in < 0) return CPB_EINVAL;
    const int64_t l = (int64_t)((double)n_in * ((double)t->k / (double)t->n));      // convcode.py:699
    if (L) *L = l;
    if (T) *T = (int64_t)((double)(l + t->M) / (double)t->k) - 1;                   // :721
    return CPB_OK;
}

int cpb_viterbi_workspace_bytes(const cpbTrellis *t, int64_t batch, int64_t n_in, int tb_depth, int mode, size_t *bytes)
{
    if (!t || !bytes || batch < 0) return CPB_EINVAL;
    int64_t L, T;
    cpb_viterbi_sizes(t, n_in, &L, &T);
    const int D = resolve_depth(t, L, tb_depth);
    const int in_dtype = (mode == CPB_VITERBI_HARD) ? CPB_U8 : CPB_F32;
    if (use_fast(t, D, mode, in_dtype)) { *bytes = 256; return CPB_OK; }
    int64_t stride;
    *bytes = generic_chunk(t, batch, T, &stride) + 256;
    return CPB_OK;
}

int cpb_viterbi_decode(const cpbTrellis *t, const void *coded_dev, int in_dtype, int64_t batch, int64_t n_in,
                       int tb_depth, int mode, uint8_t *out_bits_dev, void *workspace_dev, size_t workspace_bytes,
                       void *stream)
{
    if (!t || !coded_dev || !out_bits_dev || batch < 0 || n_in < 0) return CPB_EINVAL;
    if (mode < 0 || mode > 2) return CPB_EINVAL;        // ValueError of convcode.py:682-685
    if (in_dtype != CPB_U8 && in_dtype != CPB_F32) return CPB_EINVAL;
    if (mode != CPB_VITERBI_HARD && in_dtype != CPB_F32) return CPB_EINVAL;
    if (batch == 0) return CPB_OK;
    int64_t L, T;
    cpb_viterbi_sizes(t, n_in, &L, &T);
    const int D = resolve_depth(t, L, tb_depth);
    // the reference returns uninitialised memory when no traceback window ever closes (T < D-1) or D < 2
    if (L <= 0 || D < 2 || T < D - 1 || T > (1 << 24)) return CPB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const DeviceProps &dp = device_props();

    if (use_fast(t, D, mode, in_dtype)) {
        fast::Params p{};
        p.coded = coded_dev; p.n_in = n_in; p.batch = batch;
        p.L = (int)L; p.T = (int)T; p.D = D;
        p.R = fast::TBB + D - 2 - (t->M - 1);
        p.mode = mode; p.out = out_bits_dev;
        p.out_vec16 = ((L % 16) == 0 && (((uintptr_t)out_bits_dev) % 16) == 0) ? 1 : 0;
        Scratch ws;
        int rc = ws.acquire(workspace_dev, workspace_bytes, 256, st);
        if (rc) return rc;
        const int pack = (mode == CPB_VITERBI_HARD) ? 2 : 1;
        int bd = 32;
        if (fast::smem_bytes(p.R, bd, pack) > dp.smem_optin) { ws.release(); return CPB_EUNSUPPORTED; }
        if (pack == 1) {
            p.amax_bits = reinterpret_cast<const uint32_t *>(ws.ptr);
            cudaError_t e = cudaMemsetAsync(ws.ptr, 0, 4, st);
            if (e != cudaSuccess) { ws.release(); return record_cuda_error(e, "cudaMemsetAsync", __FILE__, __LINE__); }
            const int64_t nel = batch * n_in;
            int g = (int)std::min<int64_t>(ceil_div(nel, 256 * 8), (int64_t)dp.sm_count * 16);
            if (g < 1) g = 1;
            fast::absmax_kernel<<<g, 256, 0, st>>>(reinterpret_cast<const float *>(coded_dev), nel,
                                                    mode == CPB_VITERBI_SOFT, reinterpret_cast<uint32_t *>(ws.ptr));
        }
        if (pack == 2) {
            if (t->fast_id == 1) rc = fast::launch<Code133_171, 2>(p, bd, st);
            else if (t->fast_id == 2) rc = fast::launch<Code171_133, 2>(p, bd, st);
            else rc = fast::launch<Code5_43, 2>(p, bd, st);
        } else {
            if (t->fast_id == 1) rc = fast::launch<Code133_171, 1>(p, bd, st);
            else if (t->fast_id == 2) rc = fast::launch<Code171_133, 1>(p, bd, st);
            else rc = fast::launch<Code5_43, 1>(p, bd, st);
        }
        ws.release();
        return rc;
    }

    // generic path, chunked so the survivor scratch stays bounded
    int64_t stride = 0;
    const size_t need = generic_chunk(t, batch, T, &stride);
    Scratch ws;
    int rc = ws.acquire(workspace_dev, workspace_bytes, need, st);
    if (rc) return rc;
    const size_t smem = sizeof(float) * ((size_t)2 * t->S + (1u << t->n)) * gen::BD;
    if (smem > dp.smem_optin) { ws.release(); return CPB_EUNSUPPORTED; }
    cudaError_t e = cudaFuncSetAttribute(gen::viterbi_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ws.release(); return record_cuda_error(e, "cudaFuncSetAttribute", __FILE__, __LINE__); }
    for (int64_t f0 = 0; f0 < batch; f0 += stride) {
        gen::Params p{};
        p.coded = coded_dev; p.in_dtype = in_dtype; p.n_in = n_in; p.frame0 = f0;
        p.nframes = (int)std::min<int64_t>(stride, batch - f0);
        p.stride = stride; p.pred = t->pred_dev;
        p.k = t->k; p.n = t->n; p.S = t->S; p.I = t->I;
        p.L = (int)L; p.T = (int)T; p.D = D; p.mode = mode;
        p.winners = reinterpret_cast<uint8_t *>(ws.ptr);
        p.best = p.winners + (size_t)(T + 1) * t->S * stride;
        p.out = out_bits_dev;
        gen::viterbi_generic_kernel<<<(unsigned)ceil_div(p.nframes, gen::BD), gen::BD, smem, st>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) { ws.release(); return record_cuda_error(e, "viterbi_generic_kernel", __FILE__, __LINE__); }
    }
    ws.release();
    return CPB_OK;
}

}  // extern "C"
