// BCJR (MAP) decoding and the turbo loop on sm_100a -- replaces commpy/channelcoding/turbo.py:78-111
// (_backward_recursion), :114-158 (_forward_recursion_decoding), :163-251 (map_decode) and the iteration
// loop of :254-333 (turbo_decode).
//
// The reference works with probabilities renormalised every step; that is the exact log-MAP algorithm, so the
// kernels run it in the log domain with the exact max* (max + log1p(exp(-|a-b|))):
//   gamma_t(s,u) = -((ys_t-(2cs-1))^2 + (yp_t-(2cp-1))^2) / (2 sigma^2)         turbo.py:62-76   (cs = MSB of the
//                                                                                output symbol, cp = LSB, :97-99)
//   beta_{t-1}(s) = max*_u  beta_t(ns(s,u)) + gamma_t(s,u) + u*La_t             :106-108, beta_N = 0 (:225-226)
//   alpha_t(ns)   = max*    alpha_{t-1}(s)  + gamma_t(s,u) + u*La_t             :136-138, alpha_0 = delta(s,0)
//   L_t = La_t + max*_s[alpha_{t-1}(s)+gamma_t(s,1)+beta_t(ns(s,1))] - max*_s[... u = 0 ...]     :141-146
// (log P(u) = u*La - softplus(La); the common term cancels like the reference's normalisations do).
// Thread mapping: one LANE PER STATE, 32/S frames per warp; neighbour metrics move by warp shuffle, received
// values are loaded S steps at a time (one coalesced 4*S-byte segment per frame) and broadcast by shuffle,
// beta is parked in a global scratch [frame][t][state].  No windowing: the recursions span the whole frame
// exactly like the reference.
#include <algorithm>

#include "common.cuh"

using namespace cpb;

struct cpbTrellis;
const int32_t *cpb_trellis_next_dev(const cpbTrellis *t);
const int32_t *cpb_trellis_out_dev(const cpbTrellis *t);
void cpb_trellis_dims(const cpbTrellis *t, int *k, int *n, int *S);
const int32_t *cpb_trellis_pred_dev(const cpbTrellis *t);

namespace bcjr {

constexpr float NEG = -1.0e30f;

__device__ __forceinline__ float maxstar(float a, float b)
{
    const float m = fmaxf(a, b);
    return m + __logf(1.0f + __expf(-fabsf(a - b)));
}

// maximum over the S lanes of one frame
template <int S>
__device__ __forceinline__ float seg_max(float v)
{
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int S>
__global__ void __launch_bounds__(128) map_kernel(const float *__restrict__ sys, const float *__restrict__ par,
                                                  const float *__restrict__ La, int64_t batch, int N,
                                                  const int32_t *__restrict__ next_tab, const int32_t *__restrict__ out_tab,
                                                  const int32_t *__restrict__ pred_tab, float inv2s2, int mode,
                                                  float *__restrict__ beta, float *__restrict__ L_out,
                                                  uint8_t *__restrict__ bits_out)
{
    constexpr int FPW = 32 / S;                  // frames per warp
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int s = lane % S, sub = lane / S;
    const int lane0 = sub * S;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t f = warp * FPW + sub;
    const bool valid = f < batch;
    if (!valid) f = batch - 1;

    // trellis constants of this lane's state
    const int ns0 = __ldg(&next_tab[s * 2 + 0]), ns1 = __ldg(&next_tab[s * 2 + 1]);
    const int o0 = __ldg(&out_tab[s * 2 + 0]), o1 = __ldg(&out_tab[s * 2 + 1]);
    const float cs0 = (float)(2 * ((o0 >> 1) & 1) - 1), cp0 = (float)(2 * (o0 & 1) - 1);
    const float cs1 = (float)(2 * ((o1 >> 1) & 1) - 1), cp1 = (float)(2 * (o1 & 1) - 1);
    const int pa = __ldg(&pred_tab[s * 2 + 0]), pb = __ldg(&pred_tab[s * 2 + 1]);     // p | u<<8 | out<<16
    const int pa_s = pa & 0xff, pa_u = (pa >> 8) & 0xff, pa_o = (pa >> 16) & 0xff;
    const int pb_s = pb & 0xff, pb_u = (pb >> 8) & 0xff, pb_o = (pb >> 16) & 0xff;
    const float acs = (float)(2 * ((pa_o >> 1) & 1) - 1), acp = (float)(2 * (pa_o & 1) - 1);
    const float bcs = (float)(2 * ((pb_o >> 1) & 1) - 1), bcp = (float)(2 * (pb_o & 1) - 1);

    const float *fs = sys + f * N, *fp = par + f * N, *fl = La + f * N;
    float *fb = beta + f * (int64_t)(N + 1) * S;

    auto gamma = [&](float ys, float yp, float cs, float cp) {
        const float a = ys - cs, b = yp - cp;
        return -(a * a + b * b) * inv2s2;
    };

    // ---- backward recursion: beta_N = 0 for every state ----
    float bt = 0.0f;
    fb[(int64_t)N * S + s] = 0.0f;
    for (int t1 = N; t1 >= 1; t1 -= S) {             // steps t1, t1-1, ..., t1-S+1 (lane i holds step t1-i)
        const int tl = t1 - s;
        float vs = 0.f, vp = 0.f, vl = 0.f;
        if (tl >= 1) { vs = __ldg(fs + tl - 1); vp = __ldg(fp + tl - 1); vl = __ldg(fl + tl - 1); }
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const int t = t1 - i;
            if (t < 1) break;                         // warp-uniform
            const float ys = __shfl_sync(FULL, vs, lane0 + i), yp = __shfl_sync(FULL, vp, lane0 + i);
            const float la = __shfl_sync(FULL, vl, lane0 + i);
            const float b0 = __shfl_sync(FULL, bt, lane0 + ns0), b1 = __shfl_sync(FULL, bt, lane0 + ns1);
            float nb = maxstar(b0 + gamma(ys, yp, cs0, cp0), b1 + gamma(ys, yp, cs1, cp1) + la);
            bt = nb - seg_max<S>(nb);                 // keep the metrics bounded (the reference divides by the sum)
            fb[(int64_t)(t - 1) * S + s] = bt;
        }
    }

    // ---- forward recursion + a-posteriori LLR ----
    float at = (s == 0) ? 0.0f : NEG;                 // alpha_0 = delta(s, 0), turbo.py:220-221
    for (int t0 = 1; t0 <= N; t0 += S) {              // steps t0 .. t0+S-1 (lane i holds step t0+i)
        const int tl = t0 + s;
        float vs = 0.f, vp = 0.f, vl = 0.f;
        if (tl <= N) { vs = __ldg(fs + tl - 1); vp = __ldg(fp + tl - 1); vl = __ldg(fl + tl - 1); }
        float keep = 0.0f;
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const int t = t0 + i;
            if (t > N) break;
            const float ys = __shfl_sync(FULL, vs, lane0 + i), yp = __shfl_sync(FULL, vp, lane0 + i);
            const float la = __shfl_sync(FULL, vl, lane0 + i);
            const float bown = fb[(int64_t)t * S + s];
            const float b0 = __shfl_sync(FULL, bown, lane0 + ns0), b1 = __shfl_sync(FULL, bown, lane0 + ns1);
            float a0 = at + gamma(ys, yp, cs0, cp0) + b0;        // APP terms exclude the prior (:141-143)
            float a1 = at + gamma(ys, yp, cs1, cp1) + b1;
#pragma unroll
            for (int o = S / 2; o > 0; o >>= 1) {
                a0 = maxstar(a0, __shfl_xor_sync(FULL, a0, o));
                a1 = maxstar(a1, __shfl_xor_sync(FULL, a1, o));
            }
            const float L = la + (a1 - a0);                       // :145
            if (i == s) keep = L;
            const float xa = __shfl_sync(FULL, at, lane0 + pa_s), xb = __shfl_sync(FULL, at, lane0 + pb_s);
            float na = maxstar(xa + gamma(ys, yp, acs, acp) + (pa_u ? la : 0.0f),
                               xb + gamma(ys, yp, bcs, bcp) + (pb_u ? la : 0.0f));
            // a state whose predecessors are both unreachable keeps the sentinel instead of drifting
            at = fmaxf(na - seg_max<S>(na), NEG);
        }
        if (valid && tl <= N) {
            L_out[f * N + tl - 1] = keep;
            if (bits_out) bits_out[f * N + tl - 1] = (uint8_t)((mode == 1 && keep > 0.0f) ? 1 : 0);   // :148-152
        }
    }
}

// out[f][i] = a[f][perm[i]] - (b ? b[f][perm[i]] : 0)        interleave (interleavers.py:13-29) of an extrinsic
__global__ void __launch_bounds__(256) gather_sub_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                                         const int32_t *__restrict__ perm, int64_t batch, int N,
                                                         float *__restrict__ out)
{
    const int64_t total = batch * N;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = g / N;
        const int i = (int)(g - f * N);
        const int p = __ldg(&perm[i]);
        const float v = a[f * N + p] - (b ? b[f * N + p] : 0.0f);
        out[g] = v;
    }
}

// out[f][perm[i]] = a[f][i] - b[f][i]                          de-interleave (interleavers.py:31-47)
__global__ void __launch_bounds__(256) scatter_sub_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                                          const int32_t *__restrict__ perm, int64_t batch, int N,
                                                          float *__restrict__ out)
{
    const int64_t total = batch * N;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = g / N;
        const int i = (int)(g - f * N);
        out[f * N + __ldg(&perm[i])] = a[g] - b[g];
    }
}

__global__ void __launch_bounds__(256) scatter_bits_kernel(const uint8_t *__restrict__ a, const int32_t *__restrict__ perm,
                                                           int64_t batch, int N, uint8_t *__restrict__ out)
{
    const int64_t total = batch * N;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = g / N;
        const int i = (int)(g - f * N);
        out[f * N + __ldg(&perm[i])] = a[g];
    }
}

static int launch_map(const cpbTrellis *t, int S, const float *sys, const float *par, const float *La, int64_t batch,
                      int N, float noise_var, int mode, float *beta, float *L_out, uint8_t *bits, cudaStream_t st)
{
    const int32_t *nx = cpb_trellis_next_dev(t), *ot = cpb_trellis_out_dev(t), *pd = cpb_trellis_pred_dev(t);
    const float inv2s2 = 1.0f / (2.0f * noise_var);
    const int fpw = 32 / S;
    const int64_t warps = ceil_div(batch, fpw);
    const unsigned grid = (unsigned)ceil_div(warps, 4);
#define CPB_MAP(SS) case SS: map_kernel<SS><<<grid, 128, 0, st>>>(sys, par, La, batch, N, nx, ot, pd, inv2s2, mode, beta, L_out, bits); break;
    switch (S) {
        CPB_MAP(2) CPB_MAP(4) CPB_MAP(8) CPB_MAP(16) CPB_MAP(32)
    default: return CPB_EUNSUPPORTED;
    }
#undef CPB_MAP
    CPB_LAUNCH_CHECK();
    return CPB_OK;
}

static int check_trellis(const cpbTrellis *t, int *S)
{
    if (!t) return CPB_EINVAL;
    int k, n;
    cpb_trellis_dims(t, &k, &n, S);
    if (k != 1 || n != 2) return CPB_EINVAL;              // map_decode is written for rate-1/2 codes (turbo.py:165-166)
    if (*S != 2 && *S != 4 && *S != 8 && *S != 16 && *S != 32) return CPB_EUNSUPPORTED;
    return CPB_OK;
}

static int64_t chunk_frames(int64_t batch, int N, int S)
{
    const double per = (double)(N + 1) * S * 4.0 + 5.0 * N * 4.0 + N;
    int64_t c = (int64_t)(6.0e9 / per);
    if (c < 1) c = 1;
    return std::min<int64_t>(c, batch);
}

}  // namespace bcjr

extern "C" {

int cpb_map_decode(const cpbTrellis *t, const float *sys_dev, const float *par_dev, const float *L_int_dev,
                   int64_t batch, int64_t N, float noise_variance, int mode, float *L_out_dev, uint8_t *bits_out_dev,
                   void *stream)
{
    int S = 0;
    int rc = bcjr::check_trellis(t, &S);
    if (rc) return rc;
    if (!sys_dev || !par_dev || !L_int_dev || !L_out_dev || batch < 0 || N < 1 || N > (1 << 24) || !(noise_variance > 0.0f))
        return CPB_EINVAL;
    if (batch == 0) return CPB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t Fc = bcjr::chunk_frames(batch, (int)N, S);
    Scratch ws;
    rc = ws.acquire(nullptr, 0, (size_t)Fc * (N + 1) * S * sizeof(float), st);
    if (rc) return rc;
    for (int64_t f0 = 0; f0 < batch && rc == CPB_OK; f0 += Fc) {
        const int64_t nb = std::min<int64_t>(Fc, batch - f0);
        rc = bcjr::launch_map(t, S, sys_dev + f0 * N, par_dev + f0 * N, L_int_dev + f0 * N, nb, (int)N, noise_variance,
                              mode, reinterpret_cast<float *>(ws.ptr), L_out_dev + f0 * N,
                              bits_out_dev ? bits_out_dev + f0 * N : nullptr, st);
    }
    ws.release();
    return rc;
}

int cpb_turbo_decode(const cpbTrellis *t, const float *sys_dev, const float *par1_dev, const float *par2_dev,
                     const int32_t *perm_dev, int64_t batch, int64_t N, float noise_variance, int n_iter,
                     const float *L_int0_dev, uint8_t *bits_out_dev, void *stream)
{
    int S = 0;
    int rc = bcjr::check_trellis(t, &S);
    if (rc) return rc;
    if (!sys_dev || !par1_dev || !par2_dev || !perm_dev || !bits_out_dev || batch < 0 || N < 1 || N > (1 << 24) ||
        n_iter < 0 || !(noise_variance > 0.0f))
        return CPB_EINVAL;
    if (batch == 0) return CPB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const DeviceProps &dp = device_props();
    const int64_t Fc = bcjr::chunk_frames(batch, (int)N, S);
    const size_t nbeta = (size_t)Fc * (N + 1) * S, nvec = (size_t)Fc * N;
    Scratch ws;
    rc = ws.acquire(nullptr, 0, (nbeta + 5 * nvec) * sizeof(float) + nvec + 256, st);
    if (rc) return rc;
    float *beta = reinterpret_cast<float *>(ws.ptr);
    float *sys_i = beta + nbeta, *La1 = sys_i + nvec, *La2 = La1 + nvec, *L1 = La2 + nvec, *L2 = L1 + nvec;
    uint8_t *dec = reinterpret_cast<uint8_t *>(L2 + nvec);
    for (int64_t f0 = 0; f0 < batch && rc == CPB_OK; f0 += Fc) {
        const int64_t nb = std::min<int64_t>(Fc, batch - f0);
        const int64_t tot = nb * N;
        const unsigned eg = (unsigned)std::min<int64_t>(ceil_div(tot, 256), (int64_t)dp.sm_count * 32);
        const float *sy = sys_dev + f0 * N, *p1 = par1_dev + f0 * N, *p2 = par2_dev + f0 * N;
        cudaError_t e;
        if (L_int0_dev) e = cudaMemcpyAsync(La1, L_int0_dev + f0 * N, tot * sizeof(float), cudaMemcpyDeviceToDevice, st);
        else e = cudaMemsetAsync(La1, 0, tot * sizeof(float), st);                         // turbo.py:304-307
        if (e == cudaSuccess) e = cudaMemsetAsync(dec, 0, tot, st);
        if (e != cudaSuccess) { rc = record_cuda_error(e, "turbo init", __FILE__, __LINE__); break; }
        bcjr::gather_sub_kernel<<<eg, 256, 0, st>>>(sy, nullptr, perm_dev, nb, (int)N, sys_i);          // :310
        for (int it = 0; it < n_iter && rc == CPB_OK; ++it) {
            rc = bcjr::launch_map(t, S, sy, p1, La1, nb, (int)N, noise_variance, 0, beta, L1, nullptr, st);   // :315
            if (rc) break;
            bcjr::gather_sub_kernel<<<eg, 256, 0, st>>>(L1, La1, perm_dev, nb, (int)N, La2);            // :318-319
            const int mode = (it == n_iter - 1) ? 1 : 0;                                                // :320-323
            rc = bcjr::launch_map(t, S, sys_i, p2, La2, nb, (int)N, noise_variance, mode, beta, L2, dec, st);  // :326
            if (rc) break;
            bcjr::scatter_sub_kernel<<<eg, 256, 0, st>>>(L2, La2, perm_dev, nb, (int)N, La1);           // :328-329
        }
        if (rc) break;
        bcjr::scatter_bits_kernel<<<eg, 256, 0, st>>>(dec, perm_dev, nb, (int)N, bits_out_dev + f0 * N); // :331
        e = cudaGetLastError();
        if (e != cudaSuccess) rc = record_cuda_error(e, "turbo kernels", __FILE__, __LINE__);
    }
    ws.release();
    return rc;
}

}  // extern "C"
