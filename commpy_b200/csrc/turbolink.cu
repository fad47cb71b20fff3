// Transmit side of a turbo-coded BPSK-AWGN link, generated on the device -- SURVEY 8(f) row 4: what
// commpy/channelcoding/turbo.py:14-59 (turbo_encode) + a BPSK mapper + AWGN do per frame on the host, for a batch:
//     sys  = msg                                     (turbo.py:47-49: conv_encode(msg, trellis, 'rsc')[::2], tail cut :55)
//     par1 = parity stream of the component code over msg          (:50, :56)
//     par2 = parity stream of the component code over msg[p_array] (:52-54, :57)
//     y_x  = (2 x - 1) + sigma * N(0,1)              for x in (sys, par1, par2)
// The reference's termination='rsc' pads ZERO INPUT bits (convcode.py:505-527), which does not drive a recursive encoder
// back to state 0, and the tails are cut off again: the three streams are those of an unterminated encoder started in
// state 0 -- exactly what map_decode assumes (beta_N = 1 for every state, turbo.py:225-226).
//
// Randomness is counter based (Philox4x32-10, same keying as txlink.cu): message bit i of global frame f is bit (i & 127)
// of Philox(counter = (f_lo, f_hi, i >> 7, 0), key = seed); the noise of values 4q .. 4q+3 of stream j comes from
// Philox(counter = (f_lo, f_hi, q, 2 + j), key = seed) through Box-Muller.  Nothing depends on the batch split.
#include <vector>

#include "common.cuh"

using namespace cpb;

struct cpbTrellis;
void cpb_trellis_dims(const cpbTrellis *t, int *k, int *n, int *S);
void cpb_trellis_host_tables(const cpbTrellis *t, const int32_t **next, const int32_t **out);

namespace turbolink {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b)
{
    const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float u2 = (float)b * 2.3283064365386963e-10f;
    const float r = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    return make_float2(r * cs, r * sn);
}

struct Params {
    int64_t frames, N, first_frame;
    uint32_t seed_lo, seed_hi;
    float sigma;
    const int32_t *perm;
    uint32_t next_bits[2];    // next state of (s, u) packed 5 bits each: S <= 8 -> 16 entries x 3 bits ... kept generic below
    uint8_t next_tab[64], par_tab[64];   // S <= 32: [2 s + u]
    int S;
    uint8_t *msg;
    float *ysys, *ypar1, *ypar2;
};

// message bits: one thread per 128 bits
__global__ void __launch_bounds__(256) msg_kernel(const Params p)
{
    const int64_t blocks = (p.N + 127) >> 7;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.frames * blocks) return;
    const int64_t fl = gid / blocks, b = gid - fl * blocks;
    const uint64_t fg = (uint64_t)(p.first_frame + fl);
    const uint4 r = philox4x32_10(make_uint4((uint32_t)fg, (uint32_t)(fg >> 32), (uint32_t)b, 0u), p.seed_lo, p.seed_hi);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint8_t *m = p.msg + fl * p.N + (b << 7);
    const int64_t cnt = min((int64_t)128, p.N - (b << 7));
    for (int64_t i = 0; i < cnt; ++i) m[i] = (uint8_t)((w[i >> 5] >> (i & 31)) & 1u);
}

// one thread per (frame, component encoder): encoder 0 walks msg in natural order (and emits the systematic stream),
// encoder 1 walks msg[p_array]
__global__ void __launch_bounds__(128) encode_kernel(const Params p)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= 2 * p.frames) return;
    const int64_t fl = gid >> 1;
    const int enc = (int)(gid & 1);
    const uint64_t fg = (uint64_t)(p.first_frame + fl);
    const uint32_t f_lo = (uint32_t)fg, f_hi = (uint32_t)(fg >> 32);
    const uint8_t *m = p.msg + fl * p.N;
    float *ypar = (enc ? p.ypar2 : p.ypar1) + fl * p.N;
    float *ysys = p.ysys + fl * p.N;
    int state = 0;                                       // the encoder starts in state 0 (convcode.py:529)
    for (int64_t q = 0; q < p.N; q += 4) {
        const uint4 rp = philox4x32_10(make_uint4(f_lo, f_hi, (uint32_t)(q >> 2), 3u + (uint32_t)enc), p.seed_lo, p.seed_hi);
        const float2 a = box_muller(rp.x, rp.y), b = box_muller(rp.z, rp.w);
        const float nz[4] = {a.x, a.y, b.x, b.y};
        float ns[4] = {0.f, 0.f, 0.f, 0.f};
        if (enc == 0) {
            const uint4 rs = philox4x32_10(make_uint4(f_lo, f_hi, (uint32_t)(q >> 2), 2u), p.seed_lo, p.seed_hi);
            const float2 c = box_muller(rs.x, rs.y), d = box_muller(rs.z, rs.w);
            ns[0] = c.x; ns[1] = c.y; ns[2] = d.x; ns[3] = d.y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t t = q + i;
            if (t >= p.N) break;
            const int u = enc ? (int)m[__ldg(p.perm + t)] : (int)m[t];
            const int e = 2 * state + u;
            const int par = p.par_tab[e];
            state = p.next_tab[e];
            ypar[t] = fmaf(p.sigma, nz[i], par ? 1.0f : -1.0f);
            if (enc == 0) ysys[t] = fmaf(p.sigma, ns[i], u ? 1.0f : -1.0f);
        }
    }
}

}  // namespace turbolink

extern "C" int cpb_turbo_link_tx(const cpbTrellis *t, const int32_t *perm_dev, int64_t frames, int64_t N, uint64_t seed,
                                 int64_t first_frame, float noise_sigma, uint8_t *msg_dev, float *sys_dev, float *par1_dev,
                                 float *par2_dev, void *stream)
{
    if (!t || frames < 0 || N < 1 || first_frame < 0) return CPB_EINVAL;
    if (frames == 0) return CPB_OK;
    if (!perm_dev || !msg_dev || !sys_dev || !par1_dev || !par2_dev) return CPB_EINVAL;
    int k, n, S;
    cpb_trellis_dims(t, &k, &n, &S);
    if (k != 1 || n != 2 || S > 32) return CPB_EUNSUPPORTED;
    const int32_t *nst, *otab;
    cpb_trellis_host_tables(t, &nst, &otab);
    turbolink::Params p;
    memset(&p, 0, sizeof(p));
    for (int s = 0; s < S; ++s)
        for (int u = 0; u < 2; ++u) {
            if (((otab[s * 2 + u] >> 1) & 1) != u) return CPB_EUNSUPPORTED;      // systematic: MSB of the output symbol = input
            p.next_tab[2 * s + u] = (uint8_t)nst[s * 2 + u];
            p.par_tab[2 * s + u] = (uint8_t)(otab[s * 2 + u] & 1);
        }
    p.S = S;
    p.frames = frames; p.N = N; p.first_frame = first_frame;
    p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
    p.sigma = noise_sigma; p.perm = perm_dev;
    p.msg = msg_dev; p.ysys = sys_dev; p.ypar1 = par1_dev; p.ypar2 = par2_dev;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t blocks = (N + 127) >> 7;
    turbolink::msg_kernel<<<(unsigned)ceil_div(frames * blocks, 256), 256, 0, st>>>(p);
    turbolink::encode_kernel<<<(unsigned)ceil_div(2 * frames, 128), 128, 0, st>>>(p);
    CPB_LAUNCH_CHECK();
    return CPB_OK;
}
