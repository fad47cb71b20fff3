// Host-buffer entry points of the BCJR / turbo, LDPC and demapper rows: the same kernels behind the same chunked
//   H2D(chunk i+1)  ||  kernels(chunk i)  ||  D2H(chunk i-1)
// pipeline as cpb_viterbi_decode_host (pipeline.cu), each on the streams and staging buffers of the handle the call is
// made with.  A CommPy caller holds host arrays; these are the calls a drop-in wrapper makes for them.
#include <algorithm>

#include "handles.cuh"

using namespace cpb;

struct cpbLdpc;
struct cpbModem;
PipeCtx &cpb_ldpc_pipe(cpbLdpc *h);
void cpb_ldpc_dims(const cpbLdpc *h, int *m, int *n);
PipeCtx &cpb_modem_pipe(cpbModem *m);
void cpb_modem_info(const cpbModem *m, int *M, int *nb, const float **cst_dev);

extern "C" {

int cpb_map_decode_host(const cpbTrellis *t, const float *sys_host, const float *par_host, const float *L_int_host,
                        int64_t batch, int64_t N, float noise_variance, int mode, float *L_out_host, uint8_t *bits_out_host)
{
    if (t && batch == 0) return CPB_OK;
    if (!t || !sys_host || !par_host || !L_int_host || batch < 0 || N < 1) return CPB_EINVAL;
    const size_t row = (size_t)N * sizeof(float);
    std::vector<HostSeg> outs;
    outs.push_back({nullptr, L_out_host, row});                   // staged even when the caller does not want it back
    outs.push_back({nullptr, bits_out_host, (size_t)N});
    // outputs the caller did not ask for are computed into the staging buffer and not copied back
    std::vector<HostSeg> copy_outs;
    std::vector<int> map_idx;
    for (size_t k = 0; k < outs.size(); ++k)
        if (outs[k].out) { copy_outs.push_back(outs[k]); map_idx.push_back((int)k); }
    const bool want_L = L_out_host != nullptr, want_b = bits_out_host != nullptr;
    // L_out is always needed by the kernel: give it a scratch segment when it is not returned
    std::vector<HostSeg> ins = {{sys_host, nullptr, row}, {par_host, nullptr, row}, {L_int_host, nullptr, row}};
    if (!want_L) ins.push_back({sys_host, nullptr, row});         // an extra input-sized slot used as the L_out scratch
    return pipe_run(const_cast<cpbTrellis *>(t)->pipe, ins, copy_outs, batch, pipe_chunk(batch, 2048, 32),
                    [&](const std::vector<void *> &din, const std::vector<void *> &dout, int64_t, int64_t nb, cudaStream_t st) {
                        float *L = want_L ? static_cast<float *>(dout[0]) : static_cast<float *>(din[3]);
                        uint8_t *b = want_b ? static_cast<uint8_t *>(dout[want_L ? 1 : 0]) : nullptr;
                        return cpb_map_decode(t, static_cast<const float *>(din[0]), static_cast<const float *>(din[1]),
                                              static_cast<const float *>(din[2]), nb, N, noise_variance, mode, L, b, nullptr, 0, st);
                    });
}

int cpb_turbo_decode_host(const cpbTrellis *t, const float *sys_host, const float *par1_host, const float *par2_host,
                          const int32_t *perm_host, int64_t batch, int64_t N, float noise_variance, int n_iter,
                          const float *L_int0_host, uint8_t *bits_out_host)
{
    if (t && batch == 0) return CPB_OK;
    if (!t || !sys_host || !par1_host || !par2_host || !perm_host || !bits_out_host || batch < 0 || N < 1) return CPB_EINVAL;
    const size_t row = (size_t)N * sizeof(float);
    int32_t *perm_dev = nullptr;
    CPB_CUDA(cudaMalloc(&perm_dev, (size_t)N * sizeof(int32_t)));
    cudaError_t e = cudaMemcpy(perm_dev, perm_host, (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(perm_dev); return record_cuda_error(e, "perm upload", __FILE__, __LINE__); }
    std::vector<HostSeg> ins = {{sys_host, nullptr, row}, {par1_host, nullptr, row}, {par2_host, nullptr, row}};
    if (L_int0_host) ins.push_back({L_int0_host, nullptr, row});
    // (a chunk must still fill the GPU: 2,048 codewords x 6 windows are 384 warps)
    const int rc = pipe_run(const_cast<cpbTrellis *>(t)->pipe, ins, {{nullptr, bits_out_host, (size_t)N}}, batch,
                            pipe_chunk(batch, 2048, 32),
                            [&](const std::vector<void *> &din, const std::vector<void *> &dout, int64_t, int64_t nb, cudaStream_t st) {
                                return cpb_turbo_decode(t, static_cast<const float *>(din[0]), static_cast<const float *>(din[1]),
                                                        static_cast<const float *>(din[2]), perm_dev, nb, N, noise_variance, n_iter,
                                                        L_int0_host ? static_cast<const float *>(din[3]) : nullptr,
                                                        static_cast<uint8_t *>(dout[0]), nullptr, 0, st);
                            });
    cudaFree(perm_dev);
    return rc;
}

int cpb_ldpc_decode_host(const cpbLdpc *h, int algorithm, void *llr_host, int precision, int64_t batch, int n_iters,
                         uint8_t *dec_host, void *out_llr_host, int32_t *iters_host)
{
    if (h && batch == 0) return CPB_OK;
    if (!h || !llr_host || !dec_host || batch < 0 || (algorithm != 0 && algorithm != 1) ||
        (precision != CPB_LDPC_FP32 && precision != CPB_LDPC_FP64))
        return CPB_EINVAL;
    int m, n;
    cpb_ldpc_dims(h, &m, &n);
    const size_t esz = (precision == CPB_LDPC_FP64) ? 8 : 4;
    const size_t row = (size_t)n * esz;
    std::vector<HostSeg> outs = {{nullptr, dec_host, (size_t)n}};
    if (out_llr_host) outs.push_back({nullptr, out_llr_host, row});
    if (iters_host) outs.push_back({nullptr, iters_host, sizeof(int32_t)});
    return pipe_run(cpb_ldpc_pipe(const_cast<cpbLdpc *>(h)), {{llr_host, nullptr, row}}, outs, batch, pipe_chunk(batch, 128, 128),
                    [&](const std::vector<void *> &din, const std::vector<void *> &dout, int64_t f0, int64_t nb, cudaStream_t st) {
                        void *ol = out_llr_host ? dout[1] : nullptr;
                        int32_t *it = iters_host ? static_cast<int32_t *>(dout[out_llr_host ? 2 : 1]) : nullptr;
                        int rc = (algorithm == 0)
                            ? cpb_ldpc_minsum(h, din[0], precision, nb, n_iters, static_cast<uint8_t *>(dout[0]), ol, it, nullptr, 0, st)
                            : cpb_ldpc_sumproduct(h, din[0], precision, nb, n_iters, static_cast<uint8_t *>(dout[0]), ol, it, nullptr, 0, st);
                        if (rc) return rc;
                        // the reference clips the caller's array in place (ldpc.py:186): the kernel clipped its device copy,
                        // which goes back over the caller's rows
                        cudaError_t e = cudaMemcpyAsync(static_cast<char *>(llr_host) + (size_t)f0 * row, din[0], (size_t)nb * row,
                                                        cudaMemcpyDeviceToHost, st);
                        return e == cudaSuccess ? CPB_OK : record_cuda_error(e, "D2H (clipped llr)", __FILE__, __LINE__);
                    });
}

int cpb_demod_soft_host(const cpbModem *m, const float *y_host, int64_t n_sym, float noise_var, float *llr_host)
{
    if (m && n_sym == 0) return CPB_OK;
    if (!m || !y_host || !llr_host || n_sym < 0) return CPB_EINVAL;
    int M, nb;
    const float *cst;
    cpb_modem_info(m, &M, &nb, &cst);
    return pipe_run(cpb_modem_pipe(const_cast<cpbModem *>(m)), {{y_host, nullptr, 2 * sizeof(float)}},
                    {{nullptr, llr_host, (size_t)nb * sizeof(float)}}, n_sym, pipe_chunk(n_sym, 1 << 20, 4096),
                    [&](const std::vector<void *> &din, const std::vector<void *> &dout, int64_t, int64_t nsy, cudaStream_t st) {
                        return cpb_demod_soft(m, static_cast<const float *>(din[0]), nsy, noise_var, static_cast<float *>(dout[0]), st);
                    });
}

}  // extern "C"
