/*
 * commpy_oracle.c -- CPU restatement of CommPy's decoding hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under commpy_b200/ may import, link or
 * call this file; it exists so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py have an independent,
 * fp64, loop-for-loop restatement of the reference algorithms to check the
 * CUDA kernels against.  Parity is pinned: oracle/validate_against_reference.py
 * runs this code against the imported reference (veeresht/CommPy @ 9aecd7c)
 * on seeded inputs and writes tests/golden/*.npz; tests/test_oracle_golden.py
 * re-checks those fixtures on every run.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * the reference checkout).  The structure deliberately mirrors the reference
 * (ring buffers, per-step traceback, probability-domain BCJR, scalar min-sum)
 * rather than anything the GPU kernels do.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>
#undef I   /* the trellis code uses I for 2^k; the imaginary unit is _Complex_I below */

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_EBADARG 1
#define ORC_ETRELLIS 2   /* a state does not have exactly 2^k predecessors */
#define ORC_EALLOC 3

/* ------------------------------------------------------------------ */
/* Viterbi: commpy/channelcoding/convcode.py:561-749                   */
/* ------------------------------------------------------------------ */

/* convcode.py:575-587 _compute_branch_metrics */
static double orc_branch_metric(int mode, const double *r, int code, int n)
{
    double acc = 0.0;
    for (int j = 0; j < n; ++j) {
        int c = (code >> (n - 1 - j)) & 1;      /* dec2bitarray: MSB first (utilities.py:58-86) */
        if (mode == 0) {                        /* hard: hamming_dist(r.astype(int), c) */
            long long ri = (long long)r[j];     /* astype(int) truncates toward zero */
            acc += (double)(ri ^ (long long)c);
        } else if (mode == 1) {                 /* soft: LLR > 0 favours bit 1 */
            double n0 = log(exp(r[j]) + 1.0);
            double n1 = n0 - r[j];
            acc += c ? n1 : n0;
        } else {                                /* unquantized: squared euclid to 2c-1 */
            double d = r[j] - (double)(2 * c - 1);
            acc += d * d;
        }
    }
    return acc;
}

/*
 * One frame.  coded: len doubles.  next_state/output: S x I row-major int32.
 * decoded: L int64 (L = int(len*k/n)).  tb_depth <= 0 selects the default
 * min(5*total_memory, L) (convcode.py:701-702).  mode 0 hard, 1 soft, 2 unquantized.
 */
/* scratch of one decode; the batch entry point allocates it once and reuses it for every frame of its chunk */
typedef struct {
    int32_t *pred_state, *pred_input, *pred_cnt, *paths, *dsym;
    double *pm, *cb, *rpad;
    int64_t *dbits;
} orc_vit_ws;

static void orc_vit_ws_free(orc_vit_ws *w)
{
    free(w->pred_state); free(w->pred_input); free(w->pred_cnt); free(w->pm); free(w->paths);
    free(w->dsym); free(w->dbits); free(w->cb); free(w->rpad);
}

static int orc_viterbi_core(const double *coded, int64_t len,
                            const int32_t *next_state, const int32_t *output,
                            int k, int n, int total_memory, int S,
                            int tb_depth, int mode, int64_t *decoded, orc_vit_ws *ws);

int orc_viterbi_decode(const double *coded, int64_t len,
                       const int32_t *next_state, const int32_t *output,
                       int k, int n, int total_memory, int S,
                       int tb_depth, int mode, int64_t *decoded)
{
    orc_vit_ws ws;
    memset(&ws, 0, sizeof(ws));
    int rc = orc_viterbi_core(coded, len, next_state, output, k, n, total_memory, S, tb_depth, mode, decoded, &ws);
    orc_vit_ws_free(&ws);
    return rc;
}

static int orc_viterbi_core(const double *coded, int64_t len,
                            const int32_t *next_state, const int32_t *output,
                            int k, int n, int total_memory, int S,
                            int tb_depth, int mode, int64_t *decoded, orc_vit_ws *ws)
{
    if (k <= 0 || n <= 0 || S <= 0 || mode < 0 || mode > 2) return ORC_EBADARG;
    const int I = 1 << k;
    const double rate = (double)k / (double)n;                 /* convcode.py:695 */
    const int64_t L = (int64_t)((double)len * rate);           /* :699 */
    if (tb_depth <= 0) {
        tb_depth = 5 * total_memory;
        if ((int64_t)tb_depth > L) tb_depth = (int)L;          /* :701-702 */
    }
    if (tb_depth < 2) return ORC_EBADARG;
    const int D = tb_depth;

    /* _where_c (convcode.py:561-572): predecessors in (prev_state asc, input asc) order */
    const int64_t nbuf = ((L + D + k - 1) / k) * k + k;
    if (!ws->pm) {        /* first frame of this workspace: sizes are identical for every frame of a batch */
        ws->pred_state = (int32_t *)malloc(sizeof(int32_t) * S * I);
        ws->pred_input = (int32_t *)malloc(sizeof(int32_t) * S * I);
        ws->pred_cnt = (int32_t *)malloc(sizeof(int32_t) * S);
        ws->pm = (double *)malloc(sizeof(double) * S * 2);
        ws->paths = (int32_t *)malloc(sizeof(int32_t) * (size_t)S * D);
        ws->dsym = (int32_t *)malloc(sizeof(int32_t) * (size_t)S * D);
        ws->dbits = (int64_t *)malloc(sizeof(int64_t) * (size_t)nbuf);
        ws->cb = (double *)malloc(sizeof(double) * (len > 0 ? len : 1));
        ws->rpad = (double *)malloc(sizeof(double) * n);
    }
    int32_t *pred_state = ws->pred_state, *pred_input = ws->pred_input, *pred_cnt = ws->pred_cnt;
    double *pm = ws->pm, *cb = ws->cb, *rpad = ws->rpad;
    int32_t *paths = ws->paths, *dsym = ws->dsym;
    int64_t *dbits = ws->dbits;
    if (!pred_state || !pred_input || !pred_cnt || !pm || !paths || !dsym || !dbits || !cb || !rpad)
        return ORC_EALLOC;
    memset(pred_cnt, 0, sizeof(int32_t) * S);
    memset(paths, 0, sizeof(int32_t) * (size_t)S * D);       /* np.empty / np.zeros of convcode.py:707-711 */
    memset(dsym, 0, sizeof(int32_t) * (size_t)S * D);
    memset(dbits, 0, sizeof(int64_t) * (size_t)nbuf);
    int rc = ORC_OK;
    for (int p = 0; p < S; ++p)
        for (int u = 0; u < I; ++u) {
            int s = next_state[p * I + u];
            if (s < 0 || s >= S || pred_cnt[s] >= I) { rc = ORC_ETRELLIS; goto done; }
            pred_state[s * I + pred_cnt[s]] = p;
            pred_input[s * I + pred_cnt[s]] = u;
            pred_cnt[s]++;
        }
    for (int s = 0; s < S; ++s)
        if (pred_cnt[s] != I) { rc = ORC_ETRELLIS; goto done; }

    for (int64_t i = 0; i < len; ++i) {
        double v = coded[i];
        if (mode == 1) { if (v > 500.0) v = 500.0; if (v < -500.0) v = -500.0; }   /* :718-719 */
        cb[i] = v;
    }
    for (int s = 0; s < S; ++s) { pm[2 * s] = INFINITY; pm[2 * s + 1] = INFINITY; } /* :705 */
    pm[0] = 0.0;                                                                    /* :706 */

    int tb_count = 1;
    int64_t count = 0;
    const int64_t t_end = (int64_t)((double)(L + total_memory) / (double)k);        /* :721 */
    const double padv = (mode == 2) ? -1.0 : 0.0;                                   /* :727-732 */
    for (int j = 0; j < n; ++j) rpad[j] = padv;

    for (int64_t t = 1; t < t_end; ++t) {
        const double *r = (t <= L / k) ? (cb + (t - 1) * n) : rpad;                 /* :723-734 */
        /* _acs_traceback (convcode.py:590-657) */
        for (int s = 0; s < S; ++s) {
            double best = 0.0; int best_i = 0;
            for (int i = 0; i < I; ++i) {
                int p = pred_state[s * I + i], u = pred_input[s * I + i];
                double m = pm[2 * p] + orc_branch_metric(mode, r, output[p * I + u], n);   /* :629 */
                if (i == 0 || m < best) { best = m; best_i = i; }   /* argmin: first minimum (:637) */
            }
            pm[2 * s + 1] = best;                                                   /* :633 */
            paths[(size_t)s * D + tb_count] = pred_state[s * I + best_i];           /* :638 */
            dsym[(size_t)s * D + tb_count] = pred_input[s * I + best_i];            /* :642 */
        }
        if (t >= D - 1) {                                                           /* :644 */
            int cur = 0; double bm = pm[1];
            for (int s = 1; s < S; ++s) if (pm[2 * s + 1] < bm) { bm = pm[2 * s + 1]; cur = s; }  /* :645 */
            for (int j = D - 1; j >= 1; --j) {                                      /* :648 */
                int sym = dsym[(size_t)cur * D + j];
                int prev = paths[(size_t)cur * D + j];
                int64_t a = t - D + 1 + (int64_t)(j - 1) * k + count;               /* :653 */
                for (int b = 0; b < k; ++b)
                    if (a + b >= 0 && a + b < nbuf) dbits[a + b] = (sym >> (k - 1 - b)) & 1;
                cur = prev;
            }
            for (int s = 0; s < S; ++s) {                                           /* :656-657 */
                memmove(&paths[(size_t)s * D], &paths[(size_t)s * D + 1], sizeof(int32_t) * (D - 1));
                memmove(&dsym[(size_t)s * D], &dsym[(size_t)s * D + 1], sizeof(int32_t) * (D - 1));
            }
        }
        if (t >= D - 1) { tb_count = D - 1; count += k - 1; } else { tb_count += 1; }   /* :739-744 */
        for (int s = 0; s < S; ++s) pm[2 * s] = pm[2 * s + 1];                      /* :747 */
    }
    for (int64_t i = 0; i < L; ++i) decoded[i] = dbits[i];                          /* :749 */
done:
    return rc;
}

/* batch of independent frames, frame-major; OpenMP over frames (CPU baseline leg) */
int orc_viterbi_decode_batch(const double *coded, int64_t batch, int64_t len,
                             const int32_t *next_state, const int32_t *output,
                             int k, int n, int total_memory, int S,
                             int tb_depth, int mode, int64_t *decoded, int nthreads)
{
    const int64_t L = (int64_t)((double)len * ((double)k / (double)n));
    int rc = ORC_OK;
    (void)nthreads;   /* frames are split over host threads by oracle.py (ctypes drops the GIL) */
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    orc_vit_ws ws;
    memset(&ws, 0, sizeof(ws));
    for (int64_t b = 0; b < batch; ++b) {
        int r = orc_viterbi_core(coded + b * len, len, next_state, output, k, n, total_memory, S,
                                 tb_depth, mode, decoded + b * L, &ws);
        if (r != ORC_OK) rc = r;
    }
    orc_vit_ws_free(&ws);
    return rc;
}

/* ------------------------------------------------------------------ */
/* BCJR / turbo: commpy/channelcoding/turbo.py:62-333                  */
/* ------------------------------------------------------------------ */

/* turbo.py:62-76 */
static double orc_branch_prob(int c0, int c1, double r0, double r1, double nv)
{
    double x = r0 - (double)(2 * c0 - 1);
    double y = r1 - (double)(2 * c1 - 1);
    return exp(-(x * x + y * y) / (2.0 * nv));
}

/*
 * map_decode (turbo.py:163-251) for an n=2 trellis.  mode: 1 = 'decode', 0 = 'compute'.
 * L_out receives the value the reference calls L_ext (= L_int + log(app1/app0), :145-146).
 */
int orc_map_decode(const double *sys, const double *par, int64_t N,
                   const int32_t *next_state, const int32_t *output, int S, int I,
                   double noise_var, const double *L_int, int mode,
                   double *L_out, int64_t *bits_out)
{
    if (I != 2) return ORC_EBADARG;
    double *beta = (double *)calloc((size_t)S * (N + 1), sizeof(double));           /* :224 */
    double *gam = (double *)calloc((size_t)I * S * (N + 1), sizeof(double));        /* :229 */
    double *pri = (double *)malloc(sizeof(double) * 2 * (N > 0 ? N : 1));
    double *f0 = (double *)calloc(S, sizeof(double));
    double *f1 = (double *)calloc(S, sizeof(double));
    if (!beta || !gam || !pri || !f0 || !f1) return ORC_EALLOC;
    for (int s = 0; s < S; ++s) beta[(size_t)s * (N + 1) + N] = 1.0;               /* :225-226 */
    for (int64_t t = 0; t < N; ++t) {                                               /* :238-240 */
        pri[t] = 1.0 / (1.0 + exp(L_int[t]));
        pri[N + t] = 1.0 - pri[t];
    }
    /* _backward_recursion (turbo.py:78-111) */
    for (int64_t rt = N; rt >= 1; --rt) {
        for (int s = 0; s < S; ++s)
            for (int u = 0; u < I; ++u) {
                int ns = next_state[s * I + u];
                int code = output[s * I + u];
                int msg_bit = (code >> 1) & 1;      /* codeword_array[0] (MSB), :99 */
                int par_bit = code & 1;             /* codeword_array[1], :98 */
                double g = orc_branch_prob(msg_bit, par_bit, sys[rt - 1], par[rt - 1], noise_var);
                gam[((size_t)u * S + s) * (N + 1) + (rt - 1)] = g;                  /* :105 */
                beta[(size_t)s * (N + 1) + (rt - 1)] +=
                    (beta[(size_t)ns * (N + 1) + rt] * g * pri[(size_t)u * N + (rt - 1)]);   /* :106-108 */
            }
        double sum = 0.0;
        for (int s = 0; s < S; ++s) sum += beta[(size_t)s * (N + 1) + (rt - 1)];
        for (int s = 0; s < S; ++s) beta[(size_t)s * (N + 1) + (rt - 1)] /= sum;    /* :110-111 */
    }
    /* _forward_recursion_decoding (turbo.py:114-158) */
    f0[0] = 1.0;                                                                    /* :220-221 */
    for (int64_t t = 1; t <= N; ++t) {
        double app[2] = {0.0, 0.0};
        for (int s = 0; s < S; ++s)
            for (int u = 0; u < I; ++u) {
                int ns = next_state[s * I + u];
                double g = gam[((size_t)u * S + s) * (N + 1) + (t - 1)];
                f1[ns] += (f0[s] * g * pri[(size_t)u * N + (t - 1)]);               /* :136-138 */
                app[u] += (f0[s] * g * beta[(size_t)ns * (N + 1) + t]);             /* :141-143 */
            }
        double lappr = L_int[t - 1] + log(app[1] / app[0]);                         /* :145 */
        L_out[t - 1] = lappr;
        if (bits_out) bits_out[t - 1] = (mode == 1 && lappr > 0) ? 1 : 0;           /* :148-152 */
        double sum = 0.0;
        for (int s = 0; s < S; ++s) sum += f1[s];
        for (int s = 0; s < S; ++s) { f0[s] = f1[s] / sum; f1[s] = 0.0; }           /* :155-158 */
    }
    free(beta); free(gam); free(pri); free(f0); free(f1);
    return ORC_OK;
}

/* turbo_decode (turbo.py:254-333). perm = interleaver.p_array (interleavers.py:13-47). */
int orc_turbo_decode(const double *sys, const double *par1, const double *par2, int64_t N,
                     const int32_t *next_state, const int32_t *output, int S, int I,
                     double noise_var, int n_iter, const int64_t *perm,
                     const double *L_int0 /* nullable */, int64_t *bits_out)
{
    double *sys_i = (double *)malloc(sizeof(double) * N);
    double *La1 = (double *)malloc(sizeof(double) * N);
    double *La2 = (double *)malloc(sizeof(double) * N);
    double *L1 = (double *)malloc(sizeof(double) * N);
    double *L2 = (double *)malloc(sizeof(double) * N);
    int64_t *dec = (int64_t *)calloc(N, sizeof(int64_t));
    if (!sys_i || !La1 || !La2 || !L1 || !L2 || !dec) return ORC_EALLOC;
    int rc = ORC_OK;
    for (int64_t i = 0; i < N; ++i) {
        La1[i] = L_int0 ? L_int0[i] : 0.0;                                          /* :304-307 */
        sys_i[i] = sys[perm[i]];                                                    /* :310 interlv */
    }
    for (int it = 0; it < n_iter; ++it) {
        rc = orc_map_decode(sys, par1, N, next_state, output, S, I, noise_var, La1, 0, L1, dec);   /* :315 */
        if (rc) goto done;
        for (int64_t i = 0; i < N; ++i) La2[i] = L1[perm[i]] - La1[perm[i]];        /* :318-319 */
        int mode = (it == n_iter - 1) ? 1 : 0;                                      /* :320-323 */
        rc = orc_map_decode(sys_i, par2, N, next_state, output, S, I, noise_var, La2, mode, L2, dec);   /* :326 */
        if (rc) goto done;
        for (int64_t i = 0; i < N; ++i) La1[perm[i]] = L2[i] - La2[i];              /* :328-329 deinterlv */
    }
    for (int64_t i = 0; i < N; ++i) bits_out[perm[i]] = dec[i];                     /* :331 */
done:
    free(sys_i); free(La1); free(La2); free(L1); free(L2); free(dec);
    return rc;
}

int orc_turbo_decode_batch(const double *sys, const double *par1, const double *par2,
                           int64_t batch, int64_t N,
                           const int32_t *next_state, const int32_t *output, int S, int I,
                           double noise_var, int n_iter, const int64_t *perm,
                           int64_t *bits_out, int nthreads)
{
    int rc = ORC_OK;
    (void)nthreads;   /* frames are split over host threads by oracle.py (ctypes drops the GIL) */
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t b = 0; b < batch; ++b) {
        int r = orc_turbo_decode(sys + b * N, par1 + b * N, par2 + b * N, N, next_state, output, S, I,
                                 noise_var, n_iter, perm, NULL, bits_out + b * N);
        if (r) rc = r;
    }
    return rc;
}

/* ------------------------------------------------------------------ */
/* LDPC min-sum: commpy/channelcoding/ldpc.py:144-254 (MSA branch)     */
/* ------------------------------------------------------------------ */

static double orc_sign(double x) { return (x > 0.0) - (x < 0.0); }   /* np.sign; sign(0)=0 */

/*
 * llr: nblocks*n doubles, clipped IN PLACE to +-500 (ldpc.py:186).
 * H in CSR: row_ptr[m+1], col_idx[nnz] (ascending within a row).
 * dec: nblocks*n int8 (block-major; the Python wrapper does the order='F' reshape of :251-254).
 * out_llr: nblocks*n doubles.  iters_done (nullable): per block, iterations executed.
 */
static int orc_ldpc_bp(double *llr, int64_t nblocks, int n, int m, const int32_t *row_ptr, const int32_t *col_idx,
                      int n_iters, int spa, int8_t *dec, double *out_llr, int32_t *iters_done);

int orc_ldpc_minsum(double *llr, int64_t nblocks, int n, int m,
                    const int32_t *row_ptr, const int32_t *col_idx,
                    int n_iters, int8_t *dec, double *out_llr, int32_t *iters_done)
{
    return orc_ldpc_bp(llr, nblocks, n, m, row_ptr, col_idx, n_iters, 0, dec, out_llr, iters_done);
}

/* sum-product variant: ldpc.py:209-227 */
int orc_ldpc_sumproduct(double *llr, int64_t nblocks, int n, int m,
                        const int32_t *row_ptr, const int32_t *col_idx,
                        int n_iters, int8_t *dec, double *out_llr, int32_t *iters_done)
{
    return orc_ldpc_bp(llr, nblocks, n, m, row_ptr, col_idx, n_iters, 1, dec, out_llr, iters_done);
}

static int orc_ldpc_bp(double *llr, int64_t nblocks, int n, int m, const int32_t *row_ptr, const int32_t *col_idx,
                      int n_iters, int spa, int8_t *dec, double *out_llr, int32_t *iters_done)
{
    const int nnz = row_ptr[m];
    for (int64_t i = 0; i < nblocks * n; ++i) {                                     /* :186 */
        if (llr[i] > 500.0) llr[i] = 500.0;
        if (llr[i] < -500.0) llr[i] = -500.0;
    }
    for (int64_t i = 0; i < nblocks * n; ++i) {                                     /* :193-194 */
        dec[i] = (int8_t)(signbit(llr[i]) ? 1 : 0);
        out_llr[i] = llr[i];
    }
    double *msg = (double *)malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
    double *tot = (double *)malloc(sizeof(double) * n);
    double *tmp = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
    if (!msg || !tot || !tmp) return ORC_EALLOC;
    for (int64_t b = 0; b < nblocks; ++b) {                                         /* :197 */
        const double *lb = llr + b * n;
        int8_t *db = dec + b * n;
        double *ob = out_llr + b * n;
        for (int i = 0; i < m; ++i)                                                 /* :199 */
            for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) msg[e] = 1.0 * lb[col_idx[e]];
        int it;
        for (it = 0; it < n_iters; ++it) {                                          /* :202 */
            int ok = 1;                                                             /* :205 */
            for (int i = 0; i < m && ok; ++i) {
                int par = 0;
                for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) par ^= db[col_idx[e]];
                if (par) ok = 0;
            }
            if (ok) break;
            if (spa) {
                for (int i = 0; i < m; ++i) {                                       /* :209-227 */
                    int b0 = row_ptr[i], deg = row_ptr[i + 1] - row_ptr[i];
                    double lsum = 0.0; int negs = 0;
                    for (int j = 0; j < deg; ++j) {
                        double t = tanh(msg[b0 + j] * .5);                          /* :211-212 */
                        msg[b0 + j] = t;
                        lsum += log2(fabs(t));                                      /* real part of log2(complex t), :217-218 */
                        if (t < 0.0) negs++;
                    }
                    double prod = exp2(lsum);                                       /* :219: exp2(sum).real = +-2^Re */
                    if (negs & 1) prod = -prod;
                    for (int j = 0; j < deg; ++j) {
                        double v = (1.0 / msg[b0 + j]) * prod;                      /* :222-223 */
                        if (v > 1.0) v = 1.0;
                        if (v < -1.0) v = -1.0;                                     /* :224 (NaN passes through) */
                        v = atanh(v) * 2.0;                                         /* :225-226 */
                        if (v > 500.0) v = 500.0;
                        if (v < -500.0) v = -500.0;                                 /* :227 */
                        msg[b0 + j] = v;
                    }
                }
            } else
            for (int i = 0; i < m; ++i) {                                           /* :230-238 */
                int b0 = row_ptr[i], deg = row_ptr[i + 1] - row_ptr[i];
                for (int j = 0; j < deg; ++j) tmp[j] = msg[b0 + j];
                for (int j = 0; j < deg; ++j) {
                    double sp = 1.0, mn = INFINITY;
                    for (int q = 0; q < deg; ++q) {
                        if (q == j) continue;
                        sp *= orc_sign(tmp[q]);
                        double a = fabs(tmp[q]);
                        if (a < mn) mn = a;
                    }
                    msg[b0 + j] = sp * mn;
                }
            }
            for (int j = 0; j < n; ++j) tot[j] = 0.0;                               /* :243 (row order) */
            for (int i = 0; i < m; ++i)
                for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) tot[col_idx[e]] += msg[e];
            for (int j = 0; j < n; ++j) tmp[j] = tot[j] + lb[j];                    /* msg_sum + llr */
            for (int i = 0; i < m; ++i)                                             /* :244-245 */
                for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
                    double v = msg[e] * -1.0;
                    msg[e] = v + tmp[col_idx[e]];
                }
            for (int j = 0; j < n; ++j) {                                           /* :247-248 */
                ob[j] = tmp[j];
                db[j] = (int8_t)(signbit(ob[j]) ? 1 : 0);
            }
        }
        if (iters_done) iters_done[b] = it;
    }
    free(msg); free(tot); free(tmp);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* Demapper: commpy/modulation.py:100-141                              */
/* ------------------------------------------------------------------ */

/* y, constellation: interleaved (re, im) doubles.  out: nsym*nb doubles, MSB-first per symbol. */
int orc_demod_soft(const double *y, int64_t nsym, const double *cst, int M, double noise_var, double *out)
{
    int nb = 0;
    while ((1 << nb) < M) nb++;
    if ((1 << nb) != M) return ORC_EBADARG;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < nsym; ++i) {
        double complex cur = y[2 * i] + y[2 * i + 1] * _Complex_I;
        for (int b = 0; b < nb; ++b) {                                              /* :129 */
            double num = 0.0, den = 0.0;
            for (int kk = 0; kk < M; ++kk) {                                        /* :132-136 */
                double complex c = cst[2 * kk] + cst[2 * kk + 1] * _Complex_I;
                double a = cabs(cur - c);
                double e = exp((-(a * a)) / noise_var);
                if ((kk >> b) & 1) num += e; else den += e;
            }
            out[i * nb + nb - 1 - b] = log(num / den);                              /* :137 */
        }
    }
    return ORC_OK;
}

/* hard decision: argmin |y - c_k| (first minimum), bits MSB-first (modulation.py:121-123) */
int orc_demod_hard(const double *y, int64_t nsym, const double *cst, int M, int8_t *out)
{
    int nb = 0;
    while ((1 << nb) < M) nb++;
    if ((1 << nb) != M) return ORC_EBADARG;
    for (int64_t i = 0; i < nsym; ++i) {
        double complex cur = y[2 * i] + y[2 * i + 1] * _Complex_I;
        int best = 0; double bd = 0.0;
        for (int kk = 0; kk < M; ++kk) {
            double complex c = cst[2 * kk] + cst[2 * kk + 1] * _Complex_I;
            double a = cabs(cur - c);
            if (kk == 0 || a < bd) { bd = a; best = kk; }
        }
        for (int b = 0; b < nb; ++b) out[i * nb + b] = (int8_t)((best >> (nb - 1 - b)) & 1);
    }
    return ORC_OK;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
