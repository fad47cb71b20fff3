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
#include <cstdlib>
#include <type_traits>
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
using Code5_7 = FFCode<6, 5, 7>;             // the 64-state trellis of commpy/channelcoding/README.md:81-84 ([[5, 7]], memory 6)

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

constexpr int TBB = 24;            // windows per traceback block
constexpr int QBITS = 19;          // |quantised LLR| <= 2^19
constexpr int QMAX = 1 << QBITS;
constexpr int BD = 32;             // one warp per CTA: every synchronisation below is a __syncwarp()
// (the input prefetch distance QD, in pairs of trellis steps, is a template parameter of the kernel body: 4 by default)
#ifndef CPB_ARITH_BITS
#define CPB_ARITH_BITS 0
#endif
constexpr bool ARITH_BITS = CPB_ARITH_BITS != 0;

struct Params {
    const void *coded;
    int64_t n_in;
    int64_t batch;
    int L, T, D, R;
    int mode;                    // CPB_VITERBI_*
    const uint32_t *amax_bits;   // float input: bits of max |x| over the call (device)
    uint8_t *out;
    int out_vec16;               // 1: rows of out are 16-byte aligned
    int in_aligned;              // 1: rows of coded are 4-byte (u8) / 16-byte (f32) aligned and n_in % 4 == 0
    uint32_t keep_mask;          // ~IDX_MASK, passed at run time so the key fix-up stays one LOP3 (reg & reg | imm)
    int sm_count;
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

__device__ __forceinline__ uint32_t mad_shift(uint32_t x, int sh, uint32_t acc)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(1u << sh), "r"(acc));     // sh is a constant after unrolling
    return r;
}

// one trellis step on register-resident keys: Kn <- ACS(K, Bm); W <- survivor bits; returns min key(s)
template <class CODE, int PACK>
__device__ __forceinline__ uint32_t acs_step(const uint32_t (&K)[64], uint32_t (&Kn)[64], const uint32_t (&Bm)[4],
                                             uint32_t (&W)[2 * PACK], const uint32_t keep)
{
    using OPS = KeyOps<PACK>;
    constexpr int H = CODE::S / 2;
    constexpr int WSH = (PACK == 2) ? 4 : 5;        // states per survivor word: 16 (packed) or 32
#pragma unroll
    for (int w = 0; w < 2 * PACK; ++w) W[w] = 0;
#pragma unroll
    for (int l = 0; l < H; ++l) {
        // predecessors 2l (survivor bit 0) and 2l+1 (bit 1); new state l <- input 0, l+H <- input 1.
        // keys are metric*64 + state: the smaller key wins, ties go to predecessor 2l (convcode.py:612-642),
        // and the winner's LSB is the survivor bit.
        const uint32_t c10 = K[2 * l + 1] + Bm[CODE::out(2 * l + 1, 0)];
        const uint32_t m0 = OPS::addmin(K[2 * l], Bm[CODE::out(2 * l, 0)], c10);
        const uint32_t c11 = K[2 * l + 1] + Bm[CODE::out(2 * l + 1, 1)];
        const uint32_t m1 = OPS::addmin(K[2 * l], Bm[CODE::out(2 * l, 1)], c11);
        Kn[l] = (m0 & keep) | OPS::idx(l);
        Kn[l + H] = (m1 & keep) | OPS::idx(l + H);
        // survivor bit = LSB of the winner.  Instead of masking it out (LOP3: ALU pipe only, and the ALU pipe is the
        // busy one) it is recovered with adds the scheduler may place on either pipe: the winner's index field is
        // 2l+d, the refreshed key's is l (resp. l+H), the metric fields are equal, so per 16/32-bit lane
        //   m0 - Kn[l] - l = d        and        (m1 + (H - l)) - Kn[l+H] = d      (no borrow between lanes).
        uint32_t x0, x1;
        if (ARITH_BITS) {
            x0 = (m0 - Kn[l]) - OPS::idx(l);
            x1 = (m1 + OPS::idx(H - l)) - Kn[l + H];
        } else {
            x0 = m0 & OPS::LSB;
            x1 = m1 & OPS::LSB;
        }
        // survivor bit -> its slot in the word: one IMAD (x * 2^sh + W) on the FMA pipe
        W[l >> WSH] = mad_shift(x0, l & ((1 << WSH) - 1), W[l >> WSH]);
        W[(l + H) >> WSH] = mad_shift(x1, l & ((1 << WSH) - 1), W[(l + H) >> WSH]);
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

// shared memory of one warp-CTA
template <int PACK>
struct Smem {
    uint32_t *w;        // survivor ring [R][2*PACK][32]
    uint16_t *best;     // best states   [R][32]   (frame B in the high byte)
    uint4 *lut;         // hard-decision branch metrics [16]
    uint32_t *tasks;    // retired-path tasks [2*TASK_CAP]
    uint32_t *ntasks;   // [1]
    uint32_t *outbits;  // [32*PACK][2]
    int R;
    __device__ __forceinline__ int get_best(int slot, int col, int fi) const { return (best[slot * BD + col] >> (8 * fi)) & 63; }
    __device__ __forceinline__ int get_dec(int slot, int col, int s, int fi) const
    {
        if (PACK == 2) return (w[(slot * 4 + (s >> 4)) * BD + col] >> ((s & 15) + 16 * fi)) & 1;
        return (w[(slot * 2 + (s >> 5)) * BD + col] >> (s & 31)) & 1;
    }
    __device__ __forceinline__ int dec_slot(int slot) const { return slot == 0 ? R - 1 : slot - 1; }
};

static size_t smem_bytes(int R, int pack)
{
    size_t b = (size_t)R * 2 * pack * BD * sizeof(uint32_t);
    b += (((size_t)R * BD * sizeof(uint16_t)) + 15) & ~(size_t)15;
    b += 16 * sizeof(uint4);
    b += (size_t)2 * 384 * sizeof(uint32_t) + 16;
    b += (size_t)BD * pack * 2 * sizeof(uint32_t);
    return b;
}

// Traceback of the windows t' in (ts, te] (App. A.1-8): bit p of the frame is the input u_{p+1} read on the
// survivor path that starts at best[min(p + D - 1, T)].  A state holds the last M inputs (newest in the MSB), so a
// walk just shifts survivor bits into a register: after k look-ups from step tau0 the low M bits of `path` are the
// state at step tau0-k and bit b is u_{tau0-k-(M-1)+b}.
//   phase A  every thread walks tau = te .. ts+1 once along the current path of each of its frames; where the path
//            misses best[tau] it is retired into a task (it still owes the bits of the windows (tau, hi] it served)
//            and a new path starts at best[tau].  16 iterations, no divergence.  The path alive at ts ("closing")
//            is finished by its own thread.
//   phase B  every retired path needs D-8 more look-ups.  All tasks have that same length and any lane can run any
//            task (the ring is in shared memory), so the warp shares them evenly, two per lane at a time: the cost
//            follows the AVERAGE number of survivor-path switches per frame, not the worst lane.
// A path that served the windows (lo, hi] and has been walked down to step lo-D+8 contributes
//   (path << (lo-ts)) & bits[lo-ts, hi-ts)          (block bit j = window - ts - 1 = output bit p0 + j),
// and in the final block the path started at T also owns every later bit up to L-1.
template <int PACK, typename PT>
struct Walker {
    static constexpr int ROWB = 2 * PACK * BD * 4;      // bytes per ring slot
    const unsigned char *wbase;
    int rowoff, wrap;
    int colsh, shbase;
    PT path;
    __device__ __forceinline__ void init(const Smem<PACK> &sm, int slot, int col, int fi, PT p0)
    {
        wbase = reinterpret_cast<const unsigned char *>(sm.w);
        rowoff = slot * ROWB; wrap = sm.R * ROWB; colsh = col * 4; shbase = 16 * fi; path = p0;
    }
    __device__ __forceinline__ void step()              // survivor look-up at the current step, move one step back
    {
        const uint32_t st = (uint32_t)path & 63u;
        const int sel = (PACK == 2) ? (int)((st << 3) & 0x180u) : (int)((st << 2) & 0x80u);   // word (st>>4 | st>>5) * 128 B
        const uint32_t word = *reinterpret_cast<const uint32_t *>(wbase + rowoff + sel + colsh);
        const uint32_t sh = (PACK == 2) ? ((st & 15u) | (uint32_t)shbase) : (st & 31u);
        path = (path << 1) | (PT)((word >> sh) & 1u);
        rowoff -= ROWB;
        if (rowoff < 0) rowoff += wrap;
    }
    __device__ __forceinline__ uint32_t state() const { return (uint32_t)path & 63u; }
};

constexpr int TASK_CAP = 384;      // retired paths kept per block and warp; beyond that they are finished inline

template <class CODE, int PACK, bool FINAL>
__device__ __forceinline__ void tb_block_body(const Smem<PACK> sm, int ts, int te, int slot_te, int D, int L,
                                               uint8_t *out0, uint8_t *out1, int valid_mask, int out_vec16)
{
    using PT = typename std::conditional<FINAL, unsigned long long, uint32_t>::type;
    constexpr int M = CODE::M;
    const int lane = threadIdx.x;
    const int p0 = ts - D + 2;                          // first output bit of this block
    const int ext = (D - 2 - (M - 1) - 1 > 0) ? (D - 2 - (M - 1) - 1) : 0;     // look-ups below the window range
    const int dsh = (D - 2 - (M - 1) - 1 < 0) ? 1 : 0;  // D = M+1: the walk is one look-up deeper than needed
    auto contribution = [&](PT path, int lo, int hi) -> PT {
        PT v = (PT)(path << (lo - ts)) >> dsh;
        v &= (PT)(~(PT)0 << (lo - ts));
        if (!(FINAL && hi == te)) v &= (PT) ~(PT)(~(PT)0 << (hi - ts));
        return v;
    };
    if (lane == 0) *sm.ntasks = 0u;
    PT acc[PACK];
    int hi[PACK];
    Walker<PACK, PT> wk[PACK];
#pragma unroll
    for (int fi = 0; fi < PACK; ++fi) {
        acc[fi] = 0; hi[fi] = te;
        wk[fi].init(sm, slot_te, lane, fi, (PT)sm.get_best(slot_te, lane, fi));
        sm.outbits[(lane + BD * fi) * 2] = 0u; sm.outbits[(lane + BD * fi) * 2 + 1] = 0u;
    }
    __syncwarp();
    // ---- phase A
    int slot = slot_te;
    for (int tau = te; tau > ts; --tau) {
        const uint32_t bw = sm.best[slot * BD + lane];
#pragma unroll
        for (int fi = 0; fi < PACK; ++fi) {
            const uint32_t b = (bw >> (8 * fi)) & 63u;
            if (tau < hi[fi] && wk[fi].state() != b) {          // the path does not pass through best[tau]: retire it
                const uint32_t qi = atomicAdd(sm.ntasks, 1u);
                if (qi < (uint32_t)TASK_CAP) {
                    sm.tasks[2 * qi] = (uint32_t)lane | ((uint32_t)fi << 5) | ((uint32_t)(tau - ts) << 8) |
                                       ((uint32_t)(hi[fi] - ts) << 16);
                    sm.tasks[2 * qi + 1] = (uint32_t)wk[fi].path;      // <= 16 look-ups so far: fits 22 bits
                } else {
                    Walker<PACK, PT> t = wk[fi];
                    for (int i = 0; i < ext; ++i) t.step();
                    acc[fi] |= contribution(t.path, tau, hi[fi]);
                }
                hi[fi] = tau;
                wk[fi].path = (PT)b;
            }
            wk[fi].step();
        }
        slot = sm.dec_slot(slot);
    }
    // ---- closing paths of this thread's own frames (interleaved for ILP)
    for (int i = 0; i < ext; ++i) {
#pragma unroll
        for (int fi = 0; fi < PACK; ++fi) wk[fi].step();
    }
#pragma unroll
    for (int fi = 0; fi < PACK; ++fi) acc[fi] |= contribution(wk[fi].path, ts, hi[fi]);
    __syncwarp();
    // ---- phase B: retired paths, two per lane at a time
    int ntasks = (int)*sm.ntasks;
    if (ntasks > TASK_CAP) ntasks = TASK_CAP;
    int slot_ts = slot_te - (te - ts);
    if (slot_ts < 0) slot_ts += sm.R;
    for (int base = 0; base < ntasks; base += 2 * BD) {
        Walker<PACK, PT> tw[2];
        int lo[2], thi[2], dst[2];
        bool on[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = base + u * BD + lane;
            on[u] = i < ntasks;
            const uint32_t a = on[u] ? sm.tasks[2 * i] : 0u;
            const int col = a & 31, fi = (a >> 5) & 1;
            lo[u] = ts + (int)((a >> 8) & 255u); thi[u] = ts + (int)((a >> 16) & 255u);
            dst[u] = (col + BD * fi) * 2;
            int sl = slot_ts + (lo[u] - ts);             // the next look-up of a retired path is at step lo
            if (sl >= sm.R) sl -= sm.R;
            tw[u].init(sm, sl, col, fi, on[u] ? (PT)sm.tasks[2 * i + 1] : (PT)0);
        }
        for (int i = 0; i < ext; ++i) {
#pragma unroll
            for (int u = 0; u < 2; ++u) tw[u].step();
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!on[u]) continue;
            const PT bits = contribution(tw[u].path, lo[u], thi[u]);
            const uint32_t blo = (uint32_t)bits;
            if (blo) atomicOr(&sm.outbits[dst[u]], blo);
            if (FINAL) {
                const uint32_t bhi = (uint32_t)((unsigned long long)bits >> 32);
                if (bhi) atomicOr(&sm.outbits[dst[u] + 1], bhi);
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int fi = 0; fi < PACK; ++fi) {
        unsigned long long a64 = (unsigned long long)acc[fi] | (unsigned long long)sm.outbits[(lane + BD * fi) * 2];
        if (FINAL) a64 |= (unsigned long long)sm.outbits[(lane + BD * fi) * 2 + 1] << 32;
        if (!((valid_mask >> fi) & 1)) continue;
        uint8_t *o = (fi == 0 ? out0 : out1) + p0;
        const int nbw = te - ts;
        if (!FINAL && out_vec16 && (p0 & 15) == 0 && nbw == 16) {
            const uint32_t b16 = (uint32_t)a64 & 0xffffu;
            uint4 v;
            v.x = (((b16 >> 0) & 15u) * 0x00204081u) & 0x01010101u;
            v.y = (((b16 >> 4) & 15u) * 0x00204081u) & 0x01010101u;
            v.z = (((b16 >> 8) & 15u) * 0x00204081u) & 0x01010101u;
            v.w = (((b16 >> 12) & 15u) * 0x00204081u) & 0x01010101u;
            *reinterpret_cast<uint4 *>(o) = v;
        } else if (!FINAL && out_vec16 && (p0 & 7) == 0 && (nbw & 7) == 0) {
            // 8 decoded bits -> 8 bytes per store (rows are 16-byte aligned, p0 is a multiple of 8)
            for (int g8 = 0; g8 < nbw; g8 += 8) {
                const uint32_t b8 = (uint32_t)(a64 >> g8) & 0xffu;
                uint2 v;
                v.x = (((b8 >> 0) & 15u) * 0x00204081u) & 0x01010101u;
                v.y = (((b8 >> 4) & 15u) * 0x00204081u) & 0x01010101u;
                *reinterpret_cast<uint2 *>(o + g8) = v;
            }
        } else {
            const int cnt = FINAL ? (L - p0) : (te - ts);
            for (int i = 0; i < cnt; ++i) o[i] = (uint8_t)((a64 >> i) & 1ull);
        }
    }
    __syncwarp();
}

// the traceback is a call in the hard kernel (its 64 packed keys stay in callee-saved registers) and inlined in the
// register-capped soft kernel (CPB_TB_INLINE_SOFT)
template <class CODE, int PACK, bool FINAL>
__device__ __noinline__ void tb_block_call(const Smem<PACK> sm, int ts, int te, int slot_te, int D, int L,
                                           uint8_t *out0, uint8_t *out1, int valid_mask, int out_vec16)
{
    tb_block_body<CODE, PACK, FINAL>(sm, ts, te, slot_te, D, L, out0, out1, valid_mask, out_vec16);
}
#ifndef CPB_TB_INLINE_SOFT
#define CPB_TB_INLINE_SOFT 1
#endif
template <class CODE, int PACK, bool FINAL>
__device__ __forceinline__ void tb_block(const Smem<PACK> sm, int ts, int te, int slot_te, int D, int L,
                                         uint8_t *out0, uint8_t *out1, int valid_mask, int out_vec16)
{
    if (PACK == 1 && CPB_TB_INLINE_SOFT) tb_block_body<CODE, PACK, FINAL>(sm, ts, te, slot_te, D, L, out0, out1, valid_mask, out_vec16);
    else tb_block_call<CODE, PACK, FINAL>(sm, ts, te, slot_te, D, L, out0, out1, valid_mask, out_vec16);
}

template <class CODE, int PACK, int QD = 4>
__device__ __forceinline__ void viterbi_fast_body(const Params &p)
{
    using OPS = KeyOps<PACK>;
    constexpr int S = CODE::S, M = CODE::M;
    static_assert(S == 64, "fast path is written for 64 states");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    Smem<PACK> sm;
    {
        unsigned char *q = smem_raw;
        sm.R = p.R;
        sm.w = reinterpret_cast<uint32_t *>(q); q += (size_t)p.R * 2 * PACK * BD * sizeof(uint32_t);
        sm.best = reinterpret_cast<uint16_t *>(q); q += (((size_t)p.R * BD * sizeof(uint16_t)) + 15) & ~(size_t)15;
        sm.lut = reinterpret_cast<uint4 *>(q); q += 16 * sizeof(uint4);
        sm.tasks = reinterpret_cast<uint32_t *>(q); q += (size_t)2 * 384 * sizeof(uint32_t);
        sm.ntasks = reinterpret_cast<uint32_t *>(q); q += 16;
        sm.outbits = reinterpret_cast<uint32_t *>(q);
    }
    if (PACK == 2) {
        // hard-decision branch metrics of two frames: entry idx = r0A | r1A<<1 | r0B<<2 | r1B<<3,
        // component o = Hamming distance to output symbol o (convcode.py:579), times 64, frame B in the high half
        if (tid < 16) {
            const int a = ((tid & 1) << 1) | ((tid >> 1) & 1), b = (((tid >> 2) & 1) << 1) | ((tid >> 3) & 1);
            uint32_t e[4];
            for (int o = 0; o < 4; ++o) e[o] = ((uint32_t)__popc(o ^ a) << 6) | ((uint32_t)__popc(o ^ b) << 22);
            sm.lut[tid] = make_uint4(e[0], e[1], e[2], e[3]);
        }
        __syncwarp();
    }

    // frames of this thread (frame B = frame A + 32)
    const int64_t base = (int64_t)blockIdx.x * BD * PACK;
    int64_t fr[PACK];
    int valid_mask = 0;
    uint8_t *outp[2] = {nullptr, nullptr};
#pragma unroll
    for (int fi = 0; fi < PACK; ++fi) {
        fr[fi] = base + (int64_t)fi * BD + tid;
        if (fr[fi] < p.batch) valid_mask |= 1 << fi; else fr[fi] = p.batch - 1;
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
    const uint32_t keep = p.keep_mask;

    const unsigned char *c8 = reinterpret_cast<const unsigned char *>(p.coded);
    const float *cf = reinterpret_cast<const float *>(p.coded);
    const int npairs_in = p.L >> 1;              // pairs of steps fully covered by received data (aligned path)

    // raw received values of the pair of steps (2*pr+1, 2*pr+2)
    struct Raw { uint32_t a, b, c, d; };
    auto load_pair = [&](int pr) {
        Raw r;
        if (PACK == 2) {
            r.a = r.b = r.c = r.d = 0u;
            if (p.in_aligned) {
                if (pr < npairs_in) {
                    r.a = __ldg(reinterpret_cast<const uint32_t *>(c8 + fr[0] * p.n_in) + pr);
                    r.b = __ldg(reinterpret_cast<const uint32_t *>(c8 + fr[PACK - 1] * p.n_in) + pr);
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int tau = 2 * pr + 1 + h;
                    if (tau <= p.L) {
                        const unsigned char *qa = c8 + fr[0] * p.n_in + 2 * (int64_t)(tau - 1);
                        const unsigned char *qb = c8 + fr[PACK - 1] * p.n_in + 2 * (int64_t)(tau - 1);
                        r.a |= ((uint32_t)__ldg(qa) | ((uint32_t)__ldg(qa + 1) << 8)) << (16 * h);
                        r.b |= ((uint32_t)__ldg(qb) | ((uint32_t)__ldg(qb + 1) << 8)) << (16 * h);
                    }
                }
            }
        } else {
            const uint32_t pb = __float_as_uint(padq);
            r.a = r.b = r.c = r.d = pb;
            const float *row = cf + fr[0] * p.n_in;
            if (p.in_aligned && pr < npairs_in) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + pr);
                r.a = __float_as_uint(v.x); r.b = __float_as_uint(v.y); r.c = __float_as_uint(v.z); r.d = __float_as_uint(v.w);
            } else if (!p.in_aligned) {
                const int tau = 2 * pr + 1;
                if (tau <= p.L) { r.a = __float_as_uint(__ldg(row + 2 * (tau - 1))); r.b = __float_as_uint(__ldg(row + 2 * (tau - 1) + 1)); }
                if (tau + 1 <= p.L) { r.c = __float_as_uint(__ldg(row + 2 * tau)); r.d = __float_as_uint(__ldg(row + 2 * tau + 1)); }
            }
        }
        return r;
    };
    // the four branch metrics (times 64) of step h (0/1) of a pair
    auto make_bm = [&](const Raw &r, int h, uint32_t (&Bm)[4]) {
        if (PACK == 2) {
            const uint32_t ta = (r.a | (r.a >> 7)) >> (16 * h), tb = (r.b | (r.b >> 7)) >> (16 * h);
            const uint4 e = sm.lut[(ta & 3u) | ((tb & 3u) << 2)];
            Bm[0] = e.x; Bm[1] = e.y; Bm[2] = e.z; Bm[3] = e.w;
        } else {
            float r0 = __uint_as_float(h ? r.c : r.a), r1 = __uint_as_float(h ? r.d : r.b);
            if (p.mode == CPB_VITERBI_SOFT) {          // convcode.py:718-719
                r0 = fminf(fmaxf(r0, -500.0f), 500.0f);
                r1 = fminf(fmaxf(r1, -500.0f), 500.0f);
            }
            const int q0 = __float2int_rn(fminf(fmaxf(r0 * scale, -(float)QMAX), (float)QMAX));
            const int q1 = __float2int_rn(fminf(fmaxf(r1 * scale, -(float)QMAX), (float)QMAX));
            // -log-likelihood of code bit c given value r, up to a per-step constant (convcode.py:581-587):
            // c = 0 costs max(q,0), c = 1 costs max(-q,0)
            const uint32_t z0 = (uint32_t)max(q0, 0) << 6, o0 = (uint32_t)max(-q0, 0) << 6;
            const uint32_t z1 = (uint32_t)max(q1, 0) << 6, o1 = (uint32_t)max(-q1, 0) << 6;
            Bm[0] = z0 + z1; Bm[1] = z0 + o1; Bm[2] = o0 + z1; Bm[3] = o0 + o1;
        }
    };

    int slot = 0;
    // Traceback blocks end every TBB steps.  (Starting alternate warps half a block out of phase was measured and
    // changes nothing: warps drift apart on their own.)
    int ts_cur = p.D - 2;
    int next_te = ts_cur + TBB;
    uint32_t W[2 * PACK];

    auto finish_step = [&](int tau, uint32_t mn, uint32_t (&Kc)[64]) {
        const uint32_t bestv = (PACK == 2) ? ((mn & 63u) | (((mn >> 16) & 63u) << 8)) : (mn & 63u);
#pragma unroll
        for (int i = 0; i < 2 * PACK; ++i) sm.w[(slot * 2 * PACK + i) * BD + tid] = W[i];
        sm.best[slot * BD + tid] = (uint16_t)bestv;
        if ((tau & 15) == 0) {          // renormalise: subtract the minimum metric from every key
            const uint32_t sub = mn & ~OPS::IDX_MASK;
#pragma unroll
            for (int s = 0; s < 64; ++s) Kc[s] -= sub;
        }
        if (tau == p.T) {
            tb_block<CODE, PACK, true>(sm, ts_cur, tau, slot, p.D, p.L, outp[0], outp[PACK - 1], valid_mask, p.out_vec16);
        } else if (tau == next_te) {
            tb_block<CODE, PACK, false>(sm, ts_cur, tau, slot, p.D, p.L, outp[0], outp[PACK - 1], valid_mask, p.out_vec16);
            ts_cur = tau;
            next_te = tau + TBB;
        }
        slot = (slot + 1 == p.R) ? 0 : slot + 1;
    };

    Raw qd[QD];
#pragma unroll
    for (int i = 0; i < QD; ++i) qd[i] = load_pair(i);
    const int npairs = (p.T + 1) >> 1;
    // branch metrics are produced one step ahead of the add-compare-select that uses them, so the look-up table
    // read (hard) / float->fixed conversion (soft) of step tau+1 overlaps the ~400 instructions of step tau
    uint32_t BmA[4], BmB[4];
    make_bm(qd[0], 0, BmA);
    for (int pr = 0; pr < npairs; ++pr) {
        const Raw cur = qd[0];
#pragma unroll
        for (int i = 0; i + 1 < QD; ++i) qd[i] = qd[i + 1];
        qd[QD - 1] = load_pair(pr + QD);            // software prefetch, QD pairs of steps ahead
        const int tau = 2 * pr + 1;
        if (PACK == 2) make_bm(cur, 1, BmB);        // (the soft kernel is register-capped: it converts in place)
        uint32_t mn = acs_step<CODE, PACK>(K, Kn, BmA, W, keep);
        finish_step(tau, mn, Kn);
        if (PACK != 2) make_bm(cur, 1, BmB);
        if (PACK == 2) make_bm(qd[0], 0, BmA);
        if (tau + 1 <= p.T) {
            mn = acs_step<CODE, PACK>(Kn, K, BmB, W, keep);
            finish_step(tau + 1, mn, K);
        }
        if (PACK != 2) make_bm(qd[0], 0, BmA);
    }
}

// hard decision: two u16x2-packed frames per thread (213 registers, 7-8 warps per SM)
template <class CODE>
__global__ void __launch_bounds__(BD) viterbi_fast_kernel_hard(const Params p) { viterbi_fast_body<CODE, 2>(p); }
// soft / unquantized: one frame per thread, capped at 168 registers so that 12 warps fit an SM; the input prefetch queue
// is one pair deep here (registers are the scarce resource: 4 pairs cost 2.26 ms instead of 1.98 ms per 65,536 frames)
template <class CODE>
__global__ void __launch_bounds__(BD, 12) viterbi_fast_kernel_soft(const Params p) { viterbi_fast_body<CODE, 1, 1>(p); }
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

template <class CODE, int PACK>
static int launch(const Params &p, cudaStream_t st)
{
    const size_t smem = smem_bytes(p.R, PACK);
    void (*kern)(const Params) = (PACK == 2) ? viterbi_fast_kernel_hard<CODE> : viterbi_fast_kernel_soft<CODE>;
    CPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = ceil_div(p.batch, (int64_t)BD * PACK);
    kern<<<(unsigned)grid, BD, smem, st>>>(p);
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
    else if (code_matches<Code5_7>(*t)) t->fast_id = 4;
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
void cpb_trellis_host_tables(const cpbTrellis *t, const int32_t **next, const int32_t **out)
{
    *next = t->next_state.data();
    *out = t->output.data();
}

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
    // test hook: CPB_VITERBI_FORCE_GENERIC=1 routes every trellis through the table-driven kernel, so the two
    // independent implementations can be compared against each other at full size (tests/test_viterbi_gpu.py)
    const char *force = getenv("CPB_VITERBI_FORCE_GENERIC");
    if (force && force[0] == '1') return false;
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
    if (mode < 0 || mode > 2) return CPB_EINVAL;        // ValueError of convcode.py:682-685
    if (t && batch == 0) return CPB_OK;                 // empty tensors carry null pointers
    if (!t || !coded_dev || !out_bits_dev || batch < 0 || n_in < 0) return CPB_EINVAL;
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
        if (fast::smem_bytes(p.R, pack) > dp.smem_optin) { ws.release(); return CPB_EUNSUPPORTED; }
        p.keep_mask = (pack == 2) ? ~fast::KeyOps<2>::IDX_MASK : ~fast::KeyOps<1>::IDX_MASK;
        p.sm_count = dp.sm_count > 0 ? dp.sm_count : 148;
        {
            const size_t row = (size_t)n_in * (pack == 2 ? 1 : 4), al = (pack == 2) ? 4 : 16;
            p.in_aligned = ((n_in % 4) == 0 && (row % al) == 0 && (((uintptr_t)coded_dev) % al) == 0) ? 1 : 0;
        }
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
            if (t->fast_id == 1) rc = fast::launch<Code133_171, 2>(p, st);
            else if (t->fast_id == 2) rc = fast::launch<Code171_133, 2>(p, st);
            else if (t->fast_id == 3) rc = fast::launch<Code5_43, 2>(p, st);
            else rc = fast::launch<Code5_7, 2>(p, st);
        } else {
            if (t->fast_id == 1) rc = fast::launch<Code133_171, 1>(p, st);
            else if (t->fast_id == 2) rc = fast::launch<Code171_133, 1>(p, st);
            else if (t->fast_id == 3) rc = fast::launch<Code5_43, 1>(p, st);
            else rc = fast::launch<Code5_7, 1>(p, st);
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
