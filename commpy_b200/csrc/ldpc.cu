// LDPC min-sum belief propagation on sm_100a -- replaces the per-edge Python loop of
// commpy/channelcoding/ldpc.py:229-238 (MSA check-node update), :243-248 (variable-node update) and the
// early-stop test of :205, for a batch of blocks at once (the reference walks blocks sequentially, :197).
//
// Formulation.  Only the check-to-variable messages R (one per edge, CSR order) and the posteriors
// post_j = sum_i R_ij + llr_j are stored; the variable-to-check message the reference keeps explicitly,
// Q_ij = (tot_j + llr_j) - R_ij (:243-245), is recomputed as post_j - R_ij when the check node needs it:
// the same fp operation on the same operands, so nothing changes numerically.  Per iteration
//   check pass (one thread per check x frame group): gather post_j of the row, syndrome parity from
//        signbit(post_j) (:205 -- evaluated on the previous iteration's posteriors, i.e. BEFORE this
//        iteration, like the reference), two-minimum + sign parity, write R_ij = prod sign(others) * min|others|;
//   variable pass (one thread per variable x frame group): tot_j = sum_i R_ij in ascending check index
//        (the reference's summation order), post_j = tot_j + llr_j.  Frames whose syndrome was already zero
//        are frozen here: their posteriors stay those of the previous iteration, exactly the reference's break.
// Layout: frames are the innermost dimension of every array ([edge][frame], [variable][frame]), so each
// access is a coalesced 16-byte vector (float4 = 4 frames, double2 = 2 frames) whatever the edge index,
// and the CSR/CSC index reads are warp-uniform.  Algorithmic HBM traffic per frame-iteration: 12*E + 8*n bytes
// in fp32 (read R + write R in the check pass, read R in the variable pass, read llr + write post).
//
// CPB_LDPC_FP64 runs the same kernels in double: min-sum is only abs/min/negate/add/sub, the adds happen in
// the reference's order and no multiply exists to be contracted into an FMA, so decisions, iteration counts
// and out_llrs equal the float64 reference bit for bit.
#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace cpb;

struct cpbLdpc {
    int m, n, nnz;
    int max_row_deg, max_col_deg;
    int32_t *row_ptr = nullptr, *col_idx = nullptr;     // CSR
    int32_t *col_ptr = nullptr, *col_edge = nullptr;    // CSC: edge ids (CSR positions) of a column, ascending check
};

namespace ldpc {

template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int V = 4; struct alignas(16) type { float v[4]; }; };
template <> struct VecOf<double> { static constexpr int V = 2; struct alignas(16) type { double v[2]; }; };

// 16-byte global load that the compiler will not move (asm volatile keeps program order among such loads)
template <typename VT>
__device__ __forceinline__ VT ld16(const void *p)
{
    static_assert(sizeof(VT) == 16, "16-byte vector expected");
    uint32_t a, b, c, d;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
    VT v;
    uint32_t *w = reinterpret_cast<uint32_t *>(&v);
    w[0] = a; w[1] = b; w[2] = c; w[3] = d;
    return v;
}

struct State {
    int32_t *unsat_iter;   // [F] last iteration (1-based) at which an unsatisfied check was seen
    int32_t *done;         // [F]
    int32_t *iters;        // [F]
};

// llr[f][j] (clipped in place, ldpc.py:186) -> llrT[j][F], post[j][F]; padded frames get +1 (a codeword)
template <typename T>
__global__ void __launch_bounds__(256) load_kernel(T *__restrict__ llr, int64_t batch, int n, int64_t F,
                                                   T *__restrict__ llrT, T *__restrict__ post)
{
    __shared__ T tile[32][33];
    const int64_t f0 = (int64_t)blockIdx.y * 32;
    const int j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int64_t f = f0 + r;
        const int j = j0 + tx;
        T v = (T)1;
        if (f < batch && j < n) {
            v = llr[f * n + j];
            v = v > (T)500 ? (T)500 : (v < (T)-500 ? (T)-500 : v);
            llr[f * n + j] = v;
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r;
        const int64_t f = f0 + tx;
        if (j < n && f < F) {
            const T v = tile[tx][r];
            llrT[(int64_t)j * F + f] = v;
            post[(int64_t)j * F + f] = v;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) store_kernel(const T *__restrict__ post, int64_t batch, int n, int64_t F,
                                                    uint8_t *__restrict__ dec, T *__restrict__ out_llr)
{
    __shared__ T tile[32][33];
    const int64_t f0 = (int64_t)blockIdx.y * 32;
    const int j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r;
        const int64_t f = f0 + tx;
        tile[r][tx] = (j < n && f < F) ? post[(int64_t)j * F + f] : (T)0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t f = f0 + r;
        const int j = j0 + tx;
        if (f < batch && j < n) {
            const T v = tile[tx][r];
            dec[f * n + j] = (uint8_t)(signbit(v) ? 1 : 0);       // ldpc.py:193,248 (-0.0 -> 1)
            if (out_llr) out_llr[f * n + j] = v;
        }
    }
}

__global__ void finish_iters_kernel(const State st, int64_t batch, int n_iters, int32_t *iters_out)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f < batch) iters_out[f] = st.done[f] ? st.iters[f] : n_iters;
}

// DEGMAX > 0: rows of degree <= DEGMAX keep their variable-to-check messages in registers between the two passes
// (one read of R and post per edge); DEGMAX == 0: any degree, messages are recomputed in the second pass.
template <typename T, int DEGMAX>
__global__ void __launch_bounds__(256) cn_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                                                 int m, int64_t F, int iter, const T *__restrict__ post,
                                                 T *__restrict__ R, const State st)
{
    using VT = typename VecOf<T>::type;
    constexpr int V = VecOf<T>::V;
    constexpr int QN = (DEGMAX > 0) ? DEGMAX : 1;
    const int64_t G = F / V;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)m * G) return;
    const int i = (int)(gid / G);
    const int64_t f = (gid - (int64_t)i * G) * V;
    bool act[V];
    bool any = false;
#pragma unroll
    for (int v = 0; v < V; ++v) { act[v] = st.done[f + v] == 0; any |= act[v]; }
    if (!any) return;
    const int e0 = __ldg(&row_ptr[i]), e1 = __ldg(&row_ptr[i + 1]);
    T min1[V], min2[V];
    int arg[V], neg[V], par[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { min1[v] = (T)INFINITY; min2[v] = (T)INFINITY; arg[v] = -1; neg[v] = 0; par[v] = 0; }
    VT q[QN];
    const int deg = e1 - e0;
    if (DEGMAX > 0) {
        // issue every load of the row before the first use (edges past the row's degree re-read its last edge, so
        // no load is predicated and the memory system sees 2*DEGMAX independent 16-byte requests per thread)
        int cix[QN];
#pragma unroll
        for (int k = 0; k < QN; ++k) cix[k] = __ldg(&col_idx[min(e0 + k, e1 - 1)]);
        VT pv[QN];
#pragma unroll
        for (int k = 0; k < QN; ++k) {
            pv[k] = *reinterpret_cast<const VT *>(post + (int64_t)cix[k] * F + f);
            q[k] = *reinterpret_cast<const VT *>(R + (int64_t)min(e0 + k, e1 - 1) * F + f);
        }
#pragma unroll
        for (int k = 0; k < QN; ++k) {
            if (k < deg) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    par[v] ^= signbit(pv[k].v[v]) ? 1 : 0;
                    const T x = pv[k].v[v] - q[k].v[v];       // Q_ij = (tot_j + llr_j) - R_ij, ldpc.py:244-245
                    q[k].v[v] = x;
                    const T a = fabs(x);
                    neg[v] += (x < (T)0) ? 1 : 0;
                    // (a branch-free fmin/fmax form was tried: fewer instructions, but ptxas then sinks the post loads next
                    // to their uses and the pass gets slower -- 180 us vs 158 us at the DVB-S2 shape)
                    if (a < min1[v]) { min2[v] = min1[v]; min1[v] = a; arg[v] = k; }
                    else if (a < min2[v]) min2[v] = a;
                }
            }
        }
    } else {
        for (int e = e0; e < e1; ++e) {
            const int c = __ldg(&col_idx[e]);
            const VT p = *reinterpret_cast<const VT *>(post + (int64_t)c * F + f);
            const VT r = *reinterpret_cast<const VT *>(R + (int64_t)e * F + f);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                par[v] ^= signbit(p.v[v]) ? 1 : 0;
                const T x = p.v[v] - r.v[v];
                const T a = fabs(x);
                neg[v] += (x < (T)0) ? 1 : 0;
                if (a < min1[v]) { min2[v] = min1[v]; min1[v] = a; arg[v] = e - e0; }
                else if (a < min2[v]) min2[v] = a;
            }
        }
    }
    // benign race: every writer stores the same value; test first so ~m writers per frame do not all hit one word
#pragma unroll
    for (int v = 0; v < V; ++v)
        if (act[v] && par[v] && st.unsat_iter[f + v] != iter + 1) st.unsat_iter[f + v] = iter + 1;
    if (DEGMAX > 0) {
#pragma unroll
        for (int k = 0; k < QN; ++k) {
            const int e = e0 + k;
            if (e < e1) {
                VT r;
                bool all = true;
#pragma unroll
                for (int v = 0; v < V; ++v) all &= act[v];
                if (!all) r = *reinterpret_cast<const VT *>(R + (int64_t)e * F + f);     // keep finished frames' messages
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const T x = q[k].v[v];
                    const T mag = (k == arg[v]) ? min2[v] : min1[v];          // min over the OTHER edges (:238)
                    const int ng = neg[v] - ((x < (T)0) ? 1 : 0);
                    const T val = (ng & 1) ? -mag : mag;                      // prod of sign(others)
                    if (act[v]) r.v[v] = val;
                }
                *reinterpret_cast<VT *>(R + (int64_t)e * F + f) = r;
            }
        }
    } else {
        for (int e = e0; e < e1; ++e) {
            const int c = __ldg(&col_idx[e]);
            const VT p = *reinterpret_cast<const VT *>(post + (int64_t)c * F + f);
            VT r = *reinterpret_cast<const VT *>(R + (int64_t)e * F + f);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const T x = p.v[v] - r.v[v];
                const T mag = ((e - e0) == arg[v]) ? min2[v] : min1[v];
                const int ng = neg[v] - ((x < (T)0) ? 1 : 0);
                const T val = (ng & 1) ? -mag : mag;
                if (act[v]) r.v[v] = val;
            }
            *reinterpret_cast<VT *>(R + (int64_t)e * F + f) = r;
        }
    }
}

// Sum-product check node (ldpc.py:209-227): t = tanh(Q/2), R_ij = 2 atanh(clip((prod_row t) / t_ij, -1, 1)) clipped to
// +-500.  Same formula as the reference (product of the whole row divided by the edge's own factor); the product is
// formed directly instead of through exp2(sum(log2(complex))) so results agree to rounding (~1e-13), not bit for bit.
template <typename T>
__global__ void __launch_bounds__(256) cn_spa_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                                                     int m, int64_t F, int iter, const T *__restrict__ post,
                                                     T *__restrict__ R, const State st)
{
    using VT = typename VecOf<T>::type;
    constexpr int V = VecOf<T>::V;
    const int64_t G = F / V;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)m * G) return;
    const int i = (int)(gid / G);
    const int64_t f = (gid - (int64_t)i * G) * V;
    bool act[V];
    bool any = false;
#pragma unroll
    for (int v = 0; v < V; ++v) { act[v] = st.done[f + v] == 0; any |= act[v]; }
    if (!any) return;
    const int e0 = __ldg(&row_ptr[i]), e1 = __ldg(&row_ptr[i + 1]);
    T prod[V];
    int par[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { prod[v] = (T)1; par[v] = 0; }
    for (int e = e0; e < e1; ++e) {
        const int c = __ldg(&col_idx[e]);
        const VT p = *reinterpret_cast<const VT *>(post + (int64_t)c * F + f);
        const VT r = *reinterpret_cast<const VT *>(R + (int64_t)e * F + f);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            par[v] ^= signbit(p.v[v]) ? 1 : 0;
            prod[v] *= tanh((p.v[v] - r.v[v]) * (T)0.5);
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
        if (act[v] && par[v] && st.unsat_iter[f + v] != iter + 1) st.unsat_iter[f + v] = iter + 1;
    for (int e = e0; e < e1; ++e) {
        const int c = __ldg(&col_idx[e]);
        const VT p = *reinterpret_cast<const VT *>(post + (int64_t)c * F + f);
        VT r = *reinterpret_cast<const VT *>(R + (int64_t)e * F + f);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const T t = tanh((p.v[v] - r.v[v]) * (T)0.5);
            T x = ((T)1 / t) * prod[v];
            x = x > (T)1 ? (T)1 : (x < (T)-1 ? (T)-1 : x);           // NaN (a zero LLR in the row) passes through, as in numpy
            x = atanh(x) * (T)2;
            x = x > (T)500 ? (T)500 : (x < (T)-500 ? (T)-500 : x);
            if (act[v]) r.v[v] = x;
        }
        *reinterpret_cast<VT *>(R + (int64_t)e * F + f) = r;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) vn_kernel(const int32_t *__restrict__ col_ptr, const int32_t *__restrict__ col_edge,
                                                 int n, int64_t F, int iter, const T *__restrict__ llrT,
                                                 const T *__restrict__ R, T *__restrict__ post, const State st)
{
    using VT = typename VecOf<T>::type;
    constexpr int V = VecOf<T>::V;
    const int64_t G = F / V;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)n * G) return;
    const int j = (int)(gid / G);
    const int64_t f = (gid - (int64_t)j * G) * V;
    bool act[V];
    bool any = false;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const bool dn = st.done[f + v] != 0;
        const bool unsat = st.unsat_iter[f + v] == iter + 1;
        act[v] = !dn && unsat;
        any |= act[v];
        if (j == 0 && !dn && !unsat) {          // syndrome was zero before this iteration: the reference breaks (:205)
            st.done[f + v] = 1;
            st.iters[f + v] = iter;
        }
    }
    if (!any) return;
    const int c0 = __ldg(&col_ptr[j]), c1 = __ldg(&col_ptr[j + 1]);
    T tot[V];
#pragma unroll
    for (int v = 0; v < V; ++v) tot[v] = (T)0;
    for (int q = c0; q < c1; ++q) {             // ascending check index = the reference's summation order
        const int e = __ldg(&col_edge[q]);
        const VT r = *reinterpret_cast<const VT *>(R + (int64_t)e * F + f);
#pragma unroll
        for (int v = 0; v < V; ++v) tot[v] += r.v[v];
    }
    const VT l = *reinterpret_cast<const VT *>(llrT + (int64_t)j * F + f);
    VT p = *reinterpret_cast<const VT *>(post + (int64_t)j * F + f);
#pragma unroll
    for (int v = 0; v < V; ++v)
        if (act[v]) p.v[v] = tot[v] + l.v[v];                          // ldpc.py:247
    *reinterpret_cast<VT *>(post + (int64_t)j * F + f) = p;
}

static size_t state_bytes(int64_t F) { return (size_t)F * sizeof(int32_t) * 3; }

template <typename T>
static size_t ws_bytes(const cpbLdpc *h, int64_t F)
{
    return ((size_t)h->nnz + 2 * (size_t)h->n) * F * sizeof(T) + state_bytes(F) + 256;
}

static int64_t pad_frames(int64_t batch) { return ceil_div(batch, 32) * 32; }

// frames per pass so that the workspace stays under ~6 GB
template <typename T>
static int64_t chunk_frames(const cpbLdpc *h, int64_t batch)
{
    const double per = ((double)h->nnz + 2.0 * h->n) * sizeof(T) + 12.0;
    int64_t c = (int64_t)(6.0e9 / per);
    c = (c / 32) * 32;
    if (c < 32) c = 32;
    return std::min<int64_t>(c, pad_frames(batch));
}

template <typename T>
static int run(const cpbLdpc *h, T *llr, int64_t batch, int n_iters, int spa, uint8_t *dec, T *out_llr, int32_t *iters_out,
               void *workspace, size_t workspace_bytes, cudaStream_t st)
{
    constexpr int V = VecOf<T>::V;
    const int64_t Fc = chunk_frames<T>(h, batch);
    Scratch ws;
    int rc = ws.acquire(workspace, workspace_bytes, ws_bytes<T>(h, Fc), st);
    if (rc) return rc;
    char *base = reinterpret_cast<char *>(ws.ptr);
    T *R = reinterpret_cast<T *>(base);
    T *post = R + (size_t)h->nnz * Fc;
    T *llrT = post + (size_t)h->n * Fc;
    State s;
    s.unsat_iter = reinterpret_cast<int32_t *>(llrT + (size_t)h->n * Fc);
    s.done = s.unsat_iter + Fc;
    s.iters = s.done + Fc;
    for (int64_t f0 = 0; f0 < batch; f0 += Fc) {
        const int64_t nb = std::min<int64_t>(Fc, batch - f0);
        const int64_t F = pad_frames(nb);
        cudaError_t e = cudaMemsetAsync(R, 0, (size_t)h->nnz * F * sizeof(T), st);
        if (e == cudaSuccess) e = cudaMemsetAsync(s.unsat_iter, 0, state_bytes(Fc), st);
        if (e != cudaSuccess) { ws.release(); return record_cuda_error(e, "cudaMemsetAsync", __FILE__, __LINE__); }
        // note: R/post/llrT use stride F (<= Fc) inside this chunk
        dim3 tgrid((unsigned)ceil_div(h->n, 32), (unsigned)(F / 32));
        load_kernel<T><<<tgrid, 256, 0, st>>>(llr + f0 * h->n, nb, h->n, F, llrT, post);
        const int64_t G = F / V;
        const unsigned cn_blocks = (unsigned)ceil_div((int64_t)h->m * G, 256);
        const unsigned vn_blocks = (unsigned)ceil_div((int64_t)h->n * G, 256);
        for (int it = 0; it < n_iters; ++it) {
            if (spa) cn_spa_kernel<T><<<cn_blocks, 256, 0, st>>>(h->row_ptr, h->col_idx, h->m, F, it, post, R, s);
            else if (h->max_row_deg <= 8) cn_kernel<T, 8><<<cn_blocks, 256, 0, st>>>(h->row_ptr, h->col_idx, h->m, F, it, post, R, s);
            else cn_kernel<T, 0><<<cn_blocks, 256, 0, st>>>(h->row_ptr, h->col_idx, h->m, F, it, post, R, s);
            vn_kernel<T><<<vn_blocks, 256, 0, st>>>(h->col_ptr, h->col_edge, h->n, F, it, llrT, R, post, s);
        }
        store_kernel<T><<<tgrid, 256, 0, st>>>(post, nb, h->n, F, dec + f0 * h->n, out_llr ? out_llr + f0 * h->n : nullptr);
        if (iters_out)
            finish_iters_kernel<<<(unsigned)ceil_div(nb, 256), 256, 0, st>>>(s, nb, n_iters, iters_out + f0);
        e = cudaGetLastError();
        if (e != cudaSuccess) { ws.release(); return record_cuda_error(e, "ldpc kernels", __FILE__, __LINE__); }
    }
    ws.release();
    return CPB_OK;
}

}  // namespace ldpc

extern "C" {

int cpb_ldpc_create(const int32_t *row_ptr, const int32_t *col_idx, int m, int n, cpbLdpc **out)
{
    if (!row_ptr || !col_idx || !out || m < 1 || n < 1) return CPB_EINVAL;
    const int nnz = row_ptr[m];
    if (row_ptr[0] != 0 || nnz < 1) return CPB_EINVAL;
    cpbLdpc *h = new cpbLdpc();
    h->m = m; h->n = n; h->nnz = nnz;
    std::vector<int32_t> col_ptr(n + 1, 0), col_edge(nnz), fill(n, 0);
    int maxr = 0;
    for (int i = 0; i < m; ++i) {
        const int deg = row_ptr[i + 1] - row_ptr[i];
        if (deg < 2) { delete h; return CPB_EINVAL; }         // the reference's min over an empty set raises
        maxr = std::max(maxr, deg);
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
            if (col_idx[e] < 0 || col_idx[e] >= n) { delete h; return CPB_EINVAL; }
            col_ptr[col_idx[e] + 1]++;
        }
    }
    int maxc = 0;
    for (int j = 0; j < n; ++j) { maxc = std::max(maxc, col_ptr[j + 1]); col_ptr[j + 1] += col_ptr[j]; }
    for (int i = 0; i < m; ++i)                                // rows ascending => edges of a column in ascending check index
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
            const int j = col_idx[e];
            col_edge[col_ptr[j] + fill[j]++] = e;
        }
    h->max_row_deg = maxr; h->max_col_deg = maxc;
    cudaError_t e = cudaMalloc(&h->row_ptr, sizeof(int32_t) * (m + 1));
    if (e == cudaSuccess) e = cudaMalloc(&h->col_idx, sizeof(int32_t) * nnz);
    if (e == cudaSuccess) e = cudaMalloc(&h->col_ptr, sizeof(int32_t) * (n + 1));
    if (e == cudaSuccess) e = cudaMalloc(&h->col_edge, sizeof(int32_t) * nnz);
    if (e == cudaSuccess) e = cudaMemcpy(h->row_ptr, row_ptr, sizeof(int32_t) * (m + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->col_idx, col_idx, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->col_ptr, col_ptr.data(), sizeof(int32_t) * (n + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->col_edge, col_edge.data(), sizeof(int32_t) * nnz, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        record_cuda_error(e, "ldpc handle upload", __FILE__, __LINE__);
        cpb_ldpc_destroy(h);
        return CPB_ECUDA;
    }
    *out = h;
    return CPB_OK;
}

int cpb_ldpc_destroy(cpbLdpc *h)
{
    if (!h) return CPB_OK;
    if (h->row_ptr) cudaFree(h->row_ptr);
    if (h->col_idx) cudaFree(h->col_idx);
    if (h->col_ptr) cudaFree(h->col_ptr);
    if (h->col_edge) cudaFree(h->col_edge);
    delete h;
    return CPB_OK;
}

int cpb_ldpc_workspace_bytes(const cpbLdpc *h, int64_t batch, int precision, size_t *bytes)
{
    if (!h || !bytes || batch < 0) return CPB_EINVAL;
    if (precision == CPB_LDPC_FP64) *bytes = ldpc::ws_bytes<double>(h, ldpc::chunk_frames<double>(h, batch));
    else *bytes = ldpc::ws_bytes<float>(h, ldpc::chunk_frames<float>(h, batch));
    return CPB_OK;
}

static int ldpc_dispatch(const cpbLdpc *h, void *llr_dev, int precision, int64_t batch, int n_iters, int spa,
                         uint8_t *dec_dev, void *out_llr_dev, int32_t *iters_dev, void *workspace_dev,
                         size_t workspace_bytes, void *stream)
{
    if (h && batch == 0) return CPB_OK;
    if (!h || !llr_dev || !dec_dev || batch < 0 || n_iters < 0) return CPB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == CPB_LDPC_FP64)
        return ldpc::run<double>(h, reinterpret_cast<double *>(llr_dev), batch, n_iters, spa, dec_dev,
                                 reinterpret_cast<double *>(out_llr_dev), iters_dev, workspace_dev, workspace_bytes, st);
    if (precision == CPB_LDPC_FP32)
        return ldpc::run<float>(h, reinterpret_cast<float *>(llr_dev), batch, n_iters, spa, dec_dev,
                                reinterpret_cast<float *>(out_llr_dev), iters_dev, workspace_dev, workspace_bytes, st);
    return CPB_EINVAL;
}

int cpb_ldpc_minsum(const cpbLdpc *h, void *llr_dev, int precision, int64_t batch, int n_iters, uint8_t *dec_dev,
                    void *out_llr_dev, int32_t *iters_dev, void *workspace_dev, size_t workspace_bytes, void *stream)
{
    return ldpc_dispatch(h, llr_dev, precision, batch, n_iters, 0, dec_dev, out_llr_dev, iters_dev, workspace_dev,
                         workspace_bytes, stream);
}

int cpb_ldpc_sumproduct(const cpbLdpc *h, void *llr_dev, int precision, int64_t batch, int n_iters, uint8_t *dec_dev,
                        void *out_llr_dev, int32_t *iters_dev, void *workspace_dev, size_t workspace_bytes, void *stream)
{
    return ldpc_dispatch(h, llr_dev, precision, batch, n_iters, 1, dec_dev, out_llr_dev, iters_dev, workspace_dev,
                         workspace_bytes, stream);
}

}  // extern "C"
